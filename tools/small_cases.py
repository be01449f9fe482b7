#!/usr/bin/env python3
"""Minimal end-to-end cases of every kernel family (decode, tuned, generic, dedispersion, file runner with
time parts, corner turn), each checked against the oracle: a 2-second GPU sanity run, small enough to sit under a
memory checker where one is available.

    python tools/small_cases.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frb_baseband_b200 import _lib, spif, synth  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig, decode  # noqa: E402
from oracle import digifil_oracle as o  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(v, **kw):
    cfg = PlanConfig(out_nbit=-32, keep_bandpass=True, **kw)
    out = []
    with Plan(cfg) as pl:
        cf, fb = int(pl.chunk_frames), cfg.frame_bytes
        nfr = v[0].size // fb
        for f0 in range(0, nfr, cf):
            n = min(cf, nfr - f0)
            pl.push([x[f0 * fb:(f0 + n) * fb] for x in v])
            out.append(pl.pull().copy())
        pl.flush()
        out.append(pl.pull().copy())
        return pl.view_rows(np.concatenate(out)), pl.geometry


bw = 16.0
v = synth.make_vdif(40, seed=1, bw_mhz=bw, tone_frac=0.3, invalid_frac=0.05, fill_frac=0.05)
x, c = decode(v)
assert np.array_equal(x, o.decode_vdif(v).astype(np.float32)), "decode"
print("decode ok", c["frames_invalid"], c["frames_with_fill"])

for name, kw, okw in [
    ("tuned nchan 8 D 4", dict(nchan=8, tscrunch=4), dict(nchan=8, tscrunch_factor=4)),
    ("tuned nchan 32 coherence", dict(nchan=32, tscrunch=8, pol_mode=_lib.POL_COHERENCE), dict(nchan=32, tscrunch_factor=8, pol_mode="coherence")),
    ("generic nchan 8 L 16", dict(nchan=8, freq_res=16, chunk_units=16), dict(nchan=8, freq_res=16)),
    ("generic nchan 16 L 64", dict(nchan=16, freq_res=64, tscrunch=2, chunk_units=16), dict(nchan=16, freq_res=64, tscrunch_factor=2)),
    ("dedisp nchan 8", dict(nchan=8, tscrunch=4, dm=30.0, coherent=True), dict(nchan=8, tscrunch_factor=4, dm=30.0, coherent=True)),
]:
    rows, g = run([v], bw_mhz=[-bw], freq_mhz=[1400.0], **kw)
    if okw.get("coherent"):
        okw["nfilt"] = (int(g.nfilt_pos), int(g.nfilt_neg))
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, out_nbit=-32, keep_bandpass=True, **okw)["data"]
    if ref.size == rows.size:
        print(f"{name}: rows {rows.shape[0]}, rel err {rel(rows.reshape(ref.shape).astype(np.float64), ref.astype(np.float64)):.2e}")
    else:
        print(f"{name}: rows {rows.shape[0]} (no oracle comparison)")

# 8-bit output + rescale + splice of 2 IFs through the file runner
with tempfile.TemporaryDirectory() as d:
    paths = []
    for i in range(2):
        p = os.path.join(d, f"s_IF{i + 1}.vdif")
        synth.make_vdif(48, seed=10 + i, bw_mhz=bw).tofile(p)
        paths.append(p)
    with Plan(PlanConfig(nchan=8, bw_mhz=[-bw, bw], freq_mhz=[1400.0, 1416.0], tscrunch=4, rescale_interval_s=0.01)) as pl:
        r = pl.run_scan(paths, os.path.join(d, "o.fil"))
        print("run_scan rows", r["rows"])
        pl.run_scan(paths, None, stats_only=True)
        pl.set_rescale(*pl.rescale())
        for k in (1, 0):
            pl.run_scan(paths, os.path.join(d, "p.fil"), part=(k, 2))
    assert open(os.path.join(d, "o.fil"), "rb").read() == open(os.path.join(d, "p.fil"), "rb").read()
    print("parts ok")

# corner turn from a raw 16-bit-word recording
W, bits = spif.recipe_for_mode("VDIF_8000-1024-8-2", 4)
rng = np.random.default_rng(3)
codes = rng.integers(0, 4, size=(4, 2, 16 * 4000), dtype=np.uint8)
raw = synth.make_raw_vdif(codes, W, bits, bw_mhz=bw)
rows, _ = run([raw], nchan=8, bw_mhz=[-bw, bw, -bw, bw], freq_mhz=[1400.0 + bw * i for i in range(4)], tscrunch=4,
           raw_word_bits=W, raw_bits=bits)
print("corner turn rows", rows.shape)
print("ALL DONE")
