#!/usr/bin/env python3
"""Small seeded synthetic VDIF scans and the oracle's outputs for them, committed so that neither the oracle nor
the CUDA path can drift unnoticed ("commits small seeded synthetic VDIF + oracle outputs").
The VDIF itself is regenerated from the seed (frb_baseband_b200.synth); only the outputs are stored.

    python tests/golden/make_oracle_golden.py
"""
import base64
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from frb_baseband_b200 import synth  # noqa: E402
from oracle import digifil_oracle as o  # noqa: E402

CASES = [
    dict(name="lsb_nchan8_d64_i", nframes=40, seed=20121102, bw=16.0, usb=False, nchan=8, D=64, pol_mode="I", sig=dict(tone_frac=0.3, rho=0.2)),
    dict(name="usb_nchan32_d32_i_faults", nframes=48, seed=20121103, bw=16.0, usb=True, nchan=32, D=32, pol_mode="I",
         sig=dict(tone_frac=0.62, invalid_frac=0.05, fill_frac=0.05)),
    dict(name="lsb_nchan16_d128_coherence", nframes=64, seed=20121104, bw=32.0, usb=False, nchan=16, D=128, pol_mode="coherence",
         sig=dict(rho=0.4)),
]
out = []
for c in CASES:
    v = synth.make_vdif(c["nframes"], seed=c["seed"], bw_mhz=c["bw"], **c["sig"])
    sbw = c["bw"] if c["usb"] else -c["bw"]
    r = o.digifil(v, freq_mhz=1400.0, bw_mhz=sbw, nchan=c["nchan"], tscrunch_factor=c["D"], pol_mode=c["pol_mode"],
                  out_nbit=8, rescale_interval_s=0.002)
    rows = np.ascontiguousarray(r["data"], dtype=np.uint8)
    out.append({**{k: c[k] for k in ("name", "nframes", "seed", "bw", "usb", "nchan", "D", "pol_mode", "sig")},
                "vdif_sha256": hashlib.sha256(v.tobytes()).hexdigest(), "shape": list(rows.shape),
                "rows_b64": base64.b64encode(rows.tobytes()).decode(), "fch1": r["fch1"], "tsamp_s": r["tsamp_s"], "tstart": r["tstart"]})
json.dump({"rescale_interval_s": 0.002, "cases": out}, open(os.path.join(HERE, "oracle_small_scans.json"), "w"), indent=0)
print("wrote oracle_small_scans.json:", [(c["name"], c["shape"]) for c in out])
