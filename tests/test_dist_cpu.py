"""world_size-2 gloo test of the multi-GPU plan: clips are sharded one per rank with no data-path
collective; one all_gather reassembles the per-clip latents; the timing reduction is a MAX over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(42 + rank)  # same seeding rule as bench.py
        clip = torch.randn(1, 4, 3, 8, 8, generator=g)
        local = clip * (rank + 1)  # stand-in for the rank's own 25-step denoise
        gathered = torch.empty(world, 4, 3, 8, 8)
        dist.all_gather_into_tensor(gathered, local.contiguous())
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            torch.save({"gathered": gathered, "t": t}, out)
    finally:
        dist.destroy_process_group()


def test_clip_sharding_and_gather(tmp_path):
    out = str(tmp_path / "r0.pt")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = torch.load(out)
    for rank in range(world):
        g = torch.Generator().manual_seed(42 + rank)
        assert torch.equal(r["gathered"][rank], torch.randn(1, 4, 3, 8, 8, generator=g)[0] * (rank + 1))
    assert float(r["t"]) == 11.0
