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


class _FakePlan:
    """Stands in for a GPU plan in the CPU test of the time-sharding protocol: records what the host logic asks of it."""

    def __init__(self, rank):
        from types import SimpleNamespace
        self.nif, self.nprod, self.cfg = 2, 1, SimpleNamespace(nchan=4, keep_bandpass=False)
        self.rank, self.calls, self.preset = rank, [], None

    def rescale(self):
        assert self.rank == 0, "only the rank holding the head of the scan measures"
        m = np.arange(8, dtype=np.float32).reshape(2, 1, 4) + 0.5
        return m, 1.0 / (m + 1.0)

    def set_rescale(self, mean, scale=None):
        if mean is None:                          # back to measuring
            self.preset = None
            return
        self.preset = self.last_preset = (np.array(mean), np.array(scale))

    def run_scan(self, paths, out_path, **kw):
        self.calls.append((out_path, kw.get("part"), kw.get("stats_only", False), self.preset is not None))
        if out_path is not None:
            with open(out_path, "ab") as f:
                f.write(bytes([self.rank]))
        return {"rows": 1}


def _worker_time(rank, world, port, q, out_path):
    from frb_baseband_b200.dist import run_scan_time_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pl = _FakePlan(rank)
    pl.preset = (np.zeros(8), np.ones(8))         # left over from an earlier scan: must not reach the stats pass
    run_scan_time_sharded(pl, ["a.vdif", "b.vdif"], out_path, world, rank, nsec=3600.0)
    assert pl.preset is None                      # and the plan is handed back measuring
    q.put((rank, pl.calls, pl.last_preset[0].tolist(), pl.last_preset[1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_time_sharded_scan_protocol_world2_gloo(tmp_path):
    """Rank 0 measures the first interval without writing, everyone gets its statistics before any part is
    written, the stale output is removed once, each rank runs exactly its own part."""
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out_path = str(tmp_path / "scan.fil")
    open(out_path, "wb").write(b"stale output of an earlier run")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_time, args=(r, world, port, q, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, rest) for r, *rest in [q.get(timeout=90) for _ in range(world)])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    m = (np.arange(8, dtype=np.float32) + 0.5)
    for r in range(world):
        calls, mean, scale = got[r]
        assert np.allclose(mean, m) and np.allclose(scale, 1.0 / (m + 1.0))
        assert calls[-1] == (out_path, (r, world), False, True)              # its part, with the statistics preset
    assert got[0][0][0] == (None, None, True, False) and len(got[0][0]) == 2 and len(got[1][0]) == 1
    assert sorted(open(out_path, "rb").read()) == [0, 1]                     # stale bytes gone, both parts wrote
