#!/usr/bin/env python3
"""Small driver for ncu: a few C2-shaped chunks through the plan (device-resident input)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

nchunks = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
bws, freqs = bench.if_plan()
pl = Plan(PlanConfig(nchan=bench.NCHAN, bw_mhz=bws, freq_mhz=freqs, tscrunch=bench.TSCRUNCH, rescale_interval_s=0.2,
                     chunk_units=4))          # 4096 frames per IF and push, like bench.py
cf = int(pl.chunk_frames)
vd = [bench.make_device_vdif(torch, dev, cf * nchunks, 1 + i) for i in range(bench.NIF)]
out = torch.empty((pl.chunk_rows * (nchunks + 1), bench.NIF * bench.NCHAN), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
got = 0
for k in range(nchunks):
    pl.push([v[k * cf].data_ptr() for v in vd], nframes=cf, on_device=True)
    got += pl.pull_device(out[got].data_ptr(), out.shape[0] - got)
pl.flush()
got += pl.pull_device(out[got].data_ptr(), out.shape[0] - got)
pl.sync()
print("rows", got, "checksum", int(out[:got].to(torch.int64).sum().item()))
pl.close()
