#!/usr/bin/env python3
"""Time the column pass with parts removed (B2F_KA_VARIANT bits: 1 no butterflies, 2 no table
twiddles, 4 no shared-memory exchanges, 8 all stores into one block, 16 no global stores).  Results are wrong by construction; only the kernel
time matters.  Usage (GPU box): python tools/ablate.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for var in (0, 7, 8, 15, 16, 23):
    env = dict(os.environ, B2F_KA_VARIANT=str(var))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--seconds", "10", "--steps", "3",
                          "--cpu-sample", "0"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        k = d["kernel_ms_per_step"]
        print(f"variant {var}: column {k['column']:.2f} ms  row {k['row']:.2f} ms  step {d['ms_per_step']:.2f} ms", flush=True)
    except Exception as e:
        print("variant", var, "failed", e, out.stderr[-400:])
