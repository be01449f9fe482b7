#!/usr/bin/env python3
"""Multi-GPU check (run under torch.distributed.run with >= 2 ranks): the rows spliced by peer
stores (PeerSplice + b2f_pull_strided) equal the rows spliced by an NCCL gather."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from frb_baseband_b200.dist import PeerSplice, gather_splice, rank_if_plan  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
_, bws, freqs = rank_if_plan(bench.NIF * world, world, rank, bench.FREQ_LSB0, bench.BW)
nchunks = 3
pl = Plan(PlanConfig(nchan=bench.NCHAN, bw_mhz=bws, freq_mhz=freqs, tscrunch=bench.TSCRUNCH, rescale_interval_s=0.3, device=lr))
cf = int(pl.chunk_frames)
vd = [bench.make_device_vdif(torch, dev, cf * nchunks, 7 + 100 * rank + i) for i in range(bench.NIF)]
rows_cap = int(pl.chunk_rows) * (nchunks + 1)
tile = bench.NIF * bench.NCHAN
peer = PeerSplice(rows_cap, tile, world, rank, dev)
local = torch.zeros((rows_cap, tile), dtype=torch.uint8, device=dev)
for use_peer in (True, False):
    pl.reset()
    got = 0
    for k in range(nchunks):
        pl.push([v[k * cf].data_ptr() for v in vd], nframes=cf, on_device=True)
        got += pl.pull_strided(peer.dst(got), rows_cap - got, peer.pitch) if use_peer else pl.pull_device(local[got].data_ptr(), rows_cap - got)
    pl.flush()
    got += pl.pull_strided(peer.dst(got), rows_cap - got, peer.pitch) if use_peer else pl.pull_device(local[got].data_ptr(), rows_cap - got)
    pl.sync()
    torch.cuda.synchronize(dev)
    dist.barrier()
ref = gather_splice(local, world, rank, dst=0)
dist.barrier()
if rank == 0:
    a, b = peer.result()[:got], ref[:got]
    same = bool(torch.equal(a, b))
    print(f"peer splice vs NCCL gather on {world} GPUs: rows={got} width={a.shape[1]} equal={same} nonzero={int((a != 0).sum())}")
    assert same and got > 0
pl.close()
dist.destroy_process_group()
