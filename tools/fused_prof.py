#!/usr/bin/env python3
"""Where a warp of the fused kernel spends its cycles (B2F_FUSED_PROF=1 instrumentation, clock64 per section):
C2 shape, a few 4096-frame pushes, fused and split (two launches) side by side."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2F_FUSED_PROF"] = "1"
import torch  # noqa: E402

import bench  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

NAMES = ["col front", "wait slot", "col back", "eps/wait cols", "row load", "arrive", "row compute", "loop"]
nchunks = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
bws, freqs = bench.if_plan()
for path in ("fused", "split"):
    os.environ["B2F_PATH"] = path
    pl = Plan(PlanConfig(nchan=bench.NCHAN, bw_mhz=bws, freq_mhz=freqs, tscrunch=bench.TSCRUNCH, rescale_interval_s=0.2, chunk_units=4))
    cf = int(pl.chunk_frames)
    vd = [bench.make_device_vdif(torch, dev, cf * nchunks, 1 + i) for i in range(bench.NIF)]
    out = torch.empty((pl.chunk_rows * (nchunks + 1), bench.NIF * bench.NCHAN), dtype=torch.uint8, device=dev)
    got = 0
    for k in range(nchunks):
        pl.push([v[k * cf].data_ptr() for v in vd], nframes=cf, on_device=True)
        got += pl.pull_device(out[got].data_ptr(), out.shape[0] - got)
    pl.sync()
    prof = pl.debug(8, np.uint64).reshape(-1, 8).astype(np.float64)
    act = prof[prof.sum(axis=1) > 0]
    rounds = nchunks * 8 * 500 * 128 / len(act)        # work items per warp
    print(f"{path}: {len(act)} active warps, {rounds:.0f} rounds each; cycles per round:")
    for n, v, mx in zip(NAMES, act.mean(axis=0) / rounds, act.max(axis=0) / rounds):
        print(f"   {n:14s} {v:9.0f}   (max warp {mx:9.0f})")
    print(f"   {'total':14s} {act.sum(axis=1).mean() / rounds:9.0f}")
    pl.close()
    del vd, out
