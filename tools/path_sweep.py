#!/usr/bin/env python3
"""Round-1 kernels vs round-2 kernels (two launches) for every tuned row length: 8 IF x 32 MHz, 10 s, Stokes I, 64 us."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

dev = torch.device("cuda", 0)
bench.select_config("C2")
nif, bw, seconds = 8, 32.0, 10.0
nframes = int(seconds * 4000)
vd = [bench.make_device_vdif(torch, dev, nframes, 900 + i) for i in range(nif)]
bws = [bw if i % 2 else -bw for i in range(nif)]
for nchan in (8, 16, 32, 64, 128, 256):
    D = max(1, 2048 // nchan)                    # 64 us
    for path in ("legacy", "split"):
        os.environ["B2F_PATH"] = path
        pl = Plan(PlanConfig(nchan=nchan, bw_mhz=bws, tscrunch=D, stream=torch.cuda.current_stream(dev).cuda_stream, profile=True, chunk_units=4))
        cf = int(pl.chunk_frames)
        cap = int(seconds / pl.tsamp_s) + 2 * int(pl.chunk_rows)
        out = torch.empty((cap, int(pl.row_bytes)), dtype=torch.uint8, device=dev)

        def step():
            pl.reset()
            got = 0
            for f0 in range(0, nframes, cf):
                n = min(cf, nframes - f0)
                pl.push([v[f0].data_ptr() for v in vd], nframes=n, on_device=True)
                got += pl.pull_device(out[got].data_ptr(), cap - got)
            pl.flush()
            got += pl.pull_device(out[got].data_ptr(), cap - got)
            return got
        for _ in range(2):
            step()
        pl.sync(); torch.cuda.synchronize()
        pl.reset_timers()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            rows = step()
        pl.sync()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        kt = {k: round(v[0] / 3, 2) for k, v in pl.kernel_times().items() if v[1]}
        print(json.dumps({"nchan": nchan, "tscrunch": D, "requested": path, "path": pl.path, "ms_per_10s": round(ms, 2),
                          "rt_factor": round(seconds / (ms * 1e-3), 1), "kernel_ms": kt, "checksum": int(out[:rows].to(torch.int64).sum().item())}), flush=True)
        pl.close()
        del out
