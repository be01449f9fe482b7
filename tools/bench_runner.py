#!/usr/bin/env python3
"""File-to-file throughput of the native runner (b2f_run_scan): BASELINE config 2 split files in tmpfs ->
one spliced 8-bit filterbank.  Reports seconds of data per wall second with the files in the page cache, i.e.
the host-side ceiling of mode B (readers + pinned ring + PCIe + GPU + writer), not a disk benchmark.

    python tools/bench_runner.py [--seconds 10] [--dir /dev/shm/b2f_runner] [--ring 3]
"""
import argparse
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from frb_baseband_b200 import synth                      # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--dir", default="/dev/shm/b2f_runner")
ap.add_argument("--ring", type=int, default=3)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--readers", type=int, default=0)
a = ap.parse_args()

nif, bw, nchan, D = 8, 32.0, 128, 16
fps = 4000
os.makedirs(a.dir, exist_ok=True)
try:
    base = 1024                                           # frames generated per IF, tiled to the requested length
    nfr = int(a.seconds * fps)
    paths = []
    for i in range(nif):
        v = synth.make_vdif(base, seed=synth.config_seed(2, i), bw_mhz=bw)
        p = os.path.join(a.dir, f"c2_ef_no0001_IF{i + 1}.vdif")
        with open(p, "wb") as f:
            for k in range(0, nfr, base):
                f.write(v[: min(base, nfr - k) * 8032].tobytes())
        paths.append(p)
    bws = [bw if (i + 1) % 2 == 0 else -bw for i in range(nif)]
    freqs = [1254.0 + i * bw for i in range(nif)]
    out = os.path.join(a.dir, "c2_ef_no0001_IFall_vdif_pol2.fil")
    best = None
    with Plan(PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=D, frame_time_mode=0)) as pl:
        for rep in range(a.reps):
            t0 = time.perf_counter()
            r = pl.run_scan(paths, out, ring=a.ring, readers_per_file=a.readers)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, r)
    dt, r = best
    print(json.dumps({"workload": f"C2 files in tmpfs, {a.seconds} s", "wall_s": round(dt, 4),
                      "rt_factor": round(r["seconds_of_data"] / dt, 1), "input_GBps": round(r["bytes_in"] / dt / 1e9, 2),
                      "rows": r["rows"], "ring": a.ring, "readers_per_file": a.readers,
                      "phases_s": {k: round(r[k], 4) for k in ("setup_s", "wait_read_s", "wait_gpu_s", "write_s")}, "host_cores": len(os.sched_getaffinity(0))}))
finally:
    shutil.rmtree(a.dir, ignore_errors=True)
