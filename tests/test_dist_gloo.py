"""N>1 host logic on CPU: IF sharding and the splice gather over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frb_baseband_b200.dist import gather_splice, rank_if_plan, shard_ifs


def test_shard_ifs_partitions():
    for nif, world in [(8, 1), (8, 2), (8, 8), (16, 8), (16, 4), (6, 4)]:
        got = [i for r in range(world) for i in shard_ifs(nif, world, r)]
        assert got == list(range(1, nif + 1))
    ifs, bws, freqs = rank_if_plan(16, 2, 1, 1254.0, 32.0)
    assert ifs == list(range(9, 17)) and bws[0] == -32.0 and bws[1] == 32.0
    assert freqs[0] == 1254.0 + 8 * 32.0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, tile = 37, 16
    local = (torch.arange(rows * tile, dtype=torch.int64).reshape(rows, tile) % 200 + rank * 7).to(torch.uint8)
    out = gather_splice(local, world, rank, dst=0)
    if rank == 0:
        q.put(out.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_splice_world2_gloo():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=90)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    rows, tile = 37, 16
    base = (np.arange(rows * tile).reshape(rows, tile) % 200)
    # highest-frequency rank first
    expect = np.concatenate([(base + 7).astype(np.uint8), base.astype(np.uint8)], axis=1)
    assert np.array_equal(out, expect)
