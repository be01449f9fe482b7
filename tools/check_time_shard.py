#!/usr/bin/env python3
"""Multi-GPU check (run under torch.distributed.run with >= 2 ranks): one long scan cut into one time
segment per GPU (dist.run_scan_time_sharded: rank 0 measures the first rescale interval, broadcasts it,
every rank writes its rows at their final offset of the same file) equals the single-GPU file byte for
byte; prints both wall times.  Files live in tmpfs.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_time_shard.py [seconds]
"""
import hashlib
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from frb_baseband_b200 import synth  # noqa: E402
from frb_baseband_b200.dist import run_scan_time_sharded  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
d = "/dev/shm/b2f_time_shard"
nif, bw, fps = 8, 32.0, 4000
paths = [os.path.join(d, f"c5_ef_no0001_IF{i + 1}.vdif") for i in range(nif)]
if rank == 0:
    os.makedirs(d, exist_ok=True)
    nfr, base = int(seconds * fps), 4096
    for i, p in enumerate(paths):
        v = synth.make_vdif(base, seed=synth.config_seed(5, i), bw_mhz=bw, tone_frac=0.05 * (i + 1))
        with open(p, "wb") as f:                  # tiled 1 s pieces; header times repeat, frames are positional
            for k in range(0, nfr, base):
                f.write(v[: min(base, nfr - k) * 8032].tobytes())
dist.barrier()
try:
    bws = [bw if (i + 1) % 2 == 0 else -bw for i in range(nif)]
    freqs = [1254.0 + i * bw for i in range(nif)]
    cfg = PlanConfig(nchan=128, bw_mhz=bws, freq_mhz=freqs, tscrunch=16, device=lr)
    single, sharded = os.path.join(d, "single.fil"), os.path.join(d, "sharded.fil")
    with Plan(cfg) as pl:
        if rank == 0:
            pl.run_scan(paths, single)             # warm: pinned pool, page cache
            t0 = time.perf_counter()
            r1 = pl.run_scan(paths, single)
            t_single = time.perf_counter() - t0
        dist.barrier()
        run_scan_time_sharded(pl, paths, sharded, world, rank, device=dev)      # warm
        dist.barrier()
        t0 = time.perf_counter()
        r = run_scan_time_sharded(pl, paths, sharded, world, rank, device=dev)
        dist.barrier()
        t_shard = time.perf_counter() - t0
    rows = torch.tensor([r["rows"]], device=dev)
    dist.all_reduce(rows)
    if rank == 0:
        def digest(p):
            h = hashlib.sha256()
            with open(p, "rb") as f:
                while True:
                    b = f.read(1 << 24)
                    if not b:
                        break
                    h.update(b)
            return h.hexdigest()
        same = digest(single) == digest(sharded) and os.path.getsize(single) == os.path.getsize(sharded)
        print(json.dumps({"workload": f"8 IF x 32 MHz, {seconds:.0f} s, files in tmpfs", "gpus": world, "equal_bytes": same,
                          "rows_single": r1["rows"], "rows_sharded": int(rows.item()),
                          "single_gpu_wall_s": round(t_single, 3), "time_sharded_wall_s": round(t_shard, 3),
                          "single_rt": round(seconds / t_single, 1), "sharded_rt": round(seconds / t_shard, 1)}))
        assert same and r1["rows"] == int(rows.item())
finally:
    dist.barrier()
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    dist.destroy_process_group()
