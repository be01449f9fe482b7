#!/usr/bin/env python3
"""Device-resident throughput of the other BASELINE.json configurations (C1, C3, C4) and of the
generic path, one JSON line each.  bench.py stays on the headline workload (C2)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from frb_baseband_b200 import _lib  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
CONFIGS = {
    "C1 (4 IF x 16 MHz, nchan 32, D 32, 8-bit I, 10 s)": dict(nif=4, bw=16.0, nchan=32, D=32, seconds=10.0),
    "C1 (8 IF x 16 MHz, nchan 32, D 32, 8-bit I, 10 s)": dict(nif=8, bw=16.0, nchan=32, D=32, seconds=10.0),
    "C2 (8 IF x 32 MHz, nchan 128, D 16, 8-bit I, 20 s)": dict(nif=8, bw=32.0, nchan=128, D=16, seconds=20.0),
    "C3 (C2 with PP,QQ,Re,Im float32, 20 s)": dict(nif=8, bw=32.0, nchan=128, D=16, seconds=20.0, mode=_lib.POL_COHERENCE, nbit=-32),
    "C3b (C2 with Stokes IQUV float32, 20 s)": dict(nif=8, bw=32.0, nchan=128, D=16, seconds=20.0, mode=_lib.POL_IQUV, nbit=-32),
    "C4 (C2 + coherent dedispersion DM 560, 20 s)": dict(nif=8, bw=32.0, nchan=128, D=16, seconds=20.0, dm=560.0),
    "C5 (16 IF x 32 MHz = 4096 Mbps, nchan 128, D 16, 8-bit I, 10 s sample of the 3600 s scan)": dict(nif=16, bw=32.0, nchan=128, D=16, seconds=10.0),
    "generic (1 IF x 32 MHz, nchan 512 -> -F512:1024, D 4, 10 s)": dict(nif=1, bw=32.0, nchan=512, D=4, seconds=10.0),
    "generic policy (8 IF x 32 MHz, nchan 1024 -> -F1024:2048, D 2: submit_job.py at DM 560, 10 s)": dict(nif=8, bw=32.0, nchan=1024, D=2, seconds=10.0),
    "C4g (C4 at DM 2000: freq_res 4096 chosen from the smearing, generic kernels + kx_dedisp_generic, 10 s)": dict(nif=8, bw=32.0, nchan=128, D=16, seconds=10.0, dm=2000.0),
    "P1d (policy nchan 1024 -> -F1024:2048 with coherent dedispersion at DM 560, generic kernels + kx_dedisp_generic, 10 s)": dict(nif=8, bw=32.0, nchan=1024, D=2, seconds=10.0, dm=560.0),
}
only = sys.argv[1:] or None
for name, c in CONFIGS.items():
    if only and not any(k in name for k in only):
        continue
    nif, bw = c["nif"], c["bw"]
    fps = int(round(2 * bw * 1e6 / 16000))
    bench.FPS = fps
    nframes = int(c["seconds"] * fps)
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    freqs = [1254.0 + (i - 1) * bw for i in range(1, nif + 1)]
    nbit = c.get("nbit", 8)
    pl = Plan(PlanConfig(nchan=c["nchan"], bw_mhz=bws, freq_mhz=freqs, tscrunch=c["D"], out_nbit=nbit,
                         pol_mode=c.get("mode", _lib.POL_I), dm=c.get("dm", 0.0), coherent=c.get("dm", 0.0) > 0,
                         stream=torch.cuda.current_stream(dev).cuda_stream, profile=True))
    cf = int(pl.chunk_frames)
    vd = [bench.make_device_vdif(torch, dev, nframes, 500 + i) for i in range(nif)]
    cap = int(c["seconds"] / pl.tsamp_s) + 2 * int(pl.chunk_rows)
    out = torch.empty((cap, int(pl.row_bytes)), dtype=torch.uint8, device=dev)
    chunks = [(f0, min(cf, nframes - f0)) for f0 in range(0, nframes, cf)]

    def step():
        pl.reset()
        got = 0
        for f0, n in chunks:
            pl.push([v[f0].data_ptr() for v in vd], nframes=n, on_device=True)
            got += pl.pull_device(out[got].data_ptr(), cap - got)
        pl.flush()
        got += pl.pull_device(out[got].data_ptr(), cap - got)
        return got

    for _ in range(3):
        rows = step()
    pl.sync(); torch.cuda.synchronize()
    pl.reset_timers()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 3
    e0.record()
    for _ in range(K):
        rows = step()
    pl.sync()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    kt = {k: round(v[0] / K, 2) for k, v in pl.kernel_times().items() if v[1]}
    print(json.dumps({"workload": name, "input_GBps": round(nif * nframes * 8032 / ms / 1e6, 2), "rt_factor": round(c["seconds"] / (ms * 1e-3), 1),
                      "ms_per_step": round(ms, 2), "rows": rows, "kernel_ms": kt}), flush=True)
    pl.close()
    del vd, out
    torch.cuda.empty_cache()
