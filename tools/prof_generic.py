#!/usr/bin/env python3
"""Small driver for ncu: a few pushes of a generic-path shape (device-resident input).
usage: prof_generic.py nchan tscrunch nif nchunks"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

nchan, D, nif, nchunks = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda", 0)
bw = 32.0
bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
freqs = [1254.0 + (i - 1) * bw for i in range(1, nif + 1)]
pl = Plan(PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=D, rescale_interval_s=0.2))
cf = int(pl.chunk_frames)
vd = [bench.make_device_vdif(torch, dev, cf * nchunks, 1 + i) for i in range(nif)]
out = torch.empty((int(pl.chunk_rows) * (nchunks + 2), int(pl.row_bytes)), dtype=torch.uint8, device=dev)
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
