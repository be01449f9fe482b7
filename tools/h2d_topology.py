#!/usr/bin/env python3
"""Concurrent pinned host -> device copy rates of all ranks of one box, with and without binding each rank to the
CPU cores (and memory) of its GPU's NUMA node.  Answers whether the end-to-end scaling loss of bench.py at N >= 4
(SCALE_r01: 53.8 / 104.6 / 110.9 / 171.0 GB/s at 1 / 2 / 4 / 8 GPUs) is the box's host-memory / PCIe ceiling or
a placement problem of the ranks.  Launch:  python -m torch.distributed.run --nproc-per-node N tools/h2d_topology.py
Rank 0 prints one JSON object."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def numa_of_gpu(idx: int):
    """(pci bus id, NUMA node) of CUDA device idx; node None when sysfs does not say"""
    try:
        import subprocess
        uuid_or_idx = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[idx] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(idx)
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", uuid_or_idx], text=True).strip()
    except Exception:
        return None, None
    b = bus.lower()
    if len(b.split(":")[0]) == 8:
        b = b[4:]
    try:
        node = int(open(f"/sys/bus/pci/devices/{b}/numa_node").read())
    except Exception:
        return b, None
    return b, node


def cpus_of_node(node: int):
    out = []
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def copy_rate(dev, nbytes, reps, world):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)                       # first touch on the (possibly bound) node
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(host, non_blocking=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    gbs = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del host, d
    return gbs


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes, reps = 1 << 30, 8
    allowed = sorted(os.sched_getaffinity(0))
    bus, node = numa_of_gpu(local)
    res = {"unbound": copy_rate(dev, nbytes, reps, world)}
    bound_cpus = None
    if node is not None and node >= 0:
        cpus = [c for c in cpus_of_node(node) if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            bound_cpus = len(cpus)
            res["bound_to_gpu_numa_node"] = copy_rate(dev, nbytes, reps, world)
            os.sched_setaffinity(0, allowed)
    info = {"rank": rank, "gpu_bus": bus, "numa_node": node, "allowed_cpus": len(allowed), "bound_cpus": bound_cpus, **res}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, info)
    else:
        gathered = [info]
    if rank == 0:
        agg = {k: sum(g[k] for g in gathered if k in g) for k in ("unbound", "bound_to_gpu_numa_node")}
        print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "copies": reps, "aggregate_GBps": agg, "per_rank": gathered,
                          "host": {"cpus_visible": len(allowed), "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]) if os.path.isdir("/sys/devices/system/node") else None}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
