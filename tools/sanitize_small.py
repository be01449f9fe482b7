#!/usr/bin/env python3
"""Small scans through the three channeliser paths, for compute-sanitizer (memcheck): C2 geometry, 1 IF, 1024 frames + a
ragged second push, faulty frames included; prints a checksum per path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frb_baseband_b200 import synth  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

v = synth.make_vdif(1024 + 300, seed=5, bw_mhz=32.0, tone_frac=0.3, invalid_frac=0.01, fill_frac=0.01)
for path in ("split", "fused", "legacy"):
    os.environ["B2F_PATH"] = path
    for nchan, D in ((128, 16), (32, 8)):
        with Plan(PlanConfig(nchan=nchan, bw_mhz=[-32.0], tscrunch=D, rescale_interval_s=0.05, chunk_units=1 if nchan == 128 else 0)) as pl:
            cf = int(pl.chunk_frames)
            out = []
            for f0 in range(0, 1324, cf):
                n = min(cf, 1324 - f0)
                pl.push([v[f0 * 8032:(f0 + n) * 8032]])
                out.append(pl.pull().copy())
            pl.flush()
            out.append(pl.pull().copy())
            rows = np.concatenate([o for o in out if len(o)])
            print(path, nchan, "path", pl.path, "rows", rows.shape, "sum", int(rows.astype(np.int64).sum()), flush=True)
