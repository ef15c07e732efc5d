"""world_size-2 gloo tests of humanvid_b200.distributed: the (window x CFG-half) unit split of one clip
with its per-step all-gather, and the clip-per-rank gather of BASELINE config 4.  The UNet is replaced by a deterministic
function of the unit (the CUDA kernels need a GPU); everything else -- assignment, padding, gather, unit -> tensor map, the
window inverse map and the accumulate / CFG / DDIM arithmetic it feeds -- is the repo's own code."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from humanvid_b200.device_loop import window_inverse_map
from humanvid_b200.distributed import UnitExchange, assign_units, gather_clip_latents, unit_list
from humanvid_b200.pipeline import uniform
from humanvid_b200.scheduler import DDIMScheduler


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_unet(latents, window, half, t):
    """Stand-in for one (window, half) UNet forward: depends on the window's frames, the half and the timestep."""
    x = latents[:, :, window]
    return torch.tanh(x * (1.0 + 0.25 * half) + 0.001 * t)


def _glue_reference(latents, preds, windows, guidance, coef):
    """pipeline_pose2vid_long.py:550-563 in plain torch fp32: accumulate, /counter, CFG, DDIM (v-prediction, eta 0)."""
    un = torch.zeros_like(latents)
    tx = torch.zeros_like(latents)
    cnt = torch.zeros(1, 1, latents.shape[2], 1, 1)
    for w, win in enumerate(windows):
        un[:, :, win] += preds[(w, 0)]
        tx[:, :, win] += preds[(w, 1)]
        cnt[:, :, win] += 1
    un, tx = un / cnt, tx / cnt
    v = un + guidance * (tx - un)
    sa, sb, sp, sq = coef
    x0 = sa * latents - sb * v
    eps = sa * v + sb * latents
    return sp * x0 + sq * eps


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F = 48
        windows = list(uniform(0, 2, F, 24, 1, 4))
        units = unit_list(len(windows), True)
        ex = UnitExchange(units)
        sched = DDIMScheduler()
        sched.set_timesteps(2)
        coefs = sched.coef_table()
        g = torch.Generator().manual_seed(42)             # same seed on every rank: replicated latents
        lat = torch.randn(1, 4, F, 6, 5, generator=g)
        for step, t in enumerate(sched._host_timesteps):
            mine = [_fake_unet(lat, windows[w], h, t) for (w, h) in ex.my_units]
            preds = ex(mine)                               # the per-step collective
            lat = _glue_reference(lat, preds, windows, 3.5, [float(c) for c in coefs[step]])
        clips = gather_clip_latents(lat[:, :, :3] * (rank + 1))   # config 4: one clip per rank, one final all-gather
        torch.save({"lat": lat, "clips": clips, "mine": ex.my_units}, out + f".{rank}")
    finally:
        dist.destroy_process_group()


def test_unit_split_matches_single_process(tmp_path):
    out = str(tmp_path / "r")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = [torch.load(out + f".{k}") for k in range(world)]
    # single-process restatement of the same two steps
    F = 48
    windows = list(uniform(0, 2, F, 24, 1, 4))
    sched = DDIMScheduler()
    sched.set_timesteps(2)
    coefs = sched.coef_table()
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(1, 4, F, 6, 5, generator=g)
    for step, t in enumerate(sched._host_timesteps):
        preds = {(w, h): _fake_unet(lat, windows[w], h, t) for w in range(len(windows)) for h in (0, 1)}
        lat = _glue_reference(lat, preds, windows, 3.5, [float(c) for c in coefs[step]])
    assert torch.equal(r[0]["lat"], lat) and torch.equal(r[1]["lat"], lat)          # replicated and identical to one process
    assert r[0]["mine"] == [(0, 1), (2, 0), (2, 1)] and r[1]["mine"] == [(0, 0), (1, 0), (1, 1)]   # cost-balanced, not round-robin
    for k in range(world):
        assert torch.equal(r[0]["clips"][k], lat[0, :, :3] * (k + 1))
        assert torch.equal(r[1]["clips"], r[0]["clips"])


def test_assignment_and_inverse_map():
    units = unit_list(3, True)
    assert [len(a) for a in assign_units(units, 8)] == [1, 1, 1, 1, 1, 1, 0, 0]
    for world in (1, 2, 3, 4, 6, 8):                          # a partition of the units, balanced by cost (conditional units are heavier)
        asg = assign_units(units, world)
        assert sorted(u for a in asg for u in a) == sorted(units)
        conds = [sum(1 for u in a if u[1] == 1) for a in asg]
        assert max(conds) - min(c for c, a in zip(conds, asg) if a or world <= 6) <= 1 or world > 6
        assert max(len(a) for a in asg) == -(-len(units) // world)
    assert [len(a) for a in assign_units(units, 4)] == [2, 1, 1, 2]
    assert [sum(u[1] for u in a) for a in assign_units(units, 2)] == [2, 1]   # never all three conditional units on one rank
    ex = UnitExchange(units, world=1, rank=0)                 # degenerate world: no collective, same mapping
    got = ex([torch.full((2, 2), float(i)) for i in range(6)])
    assert all(float(got[u][0, 0]) == i for i, u in enumerate(units))
    windows = list(uniform(0, 25, 48, 24, 1, 4))
    inv = window_inverse_map(windows, 48)
    # counter pattern of SURVEY 8c(5): frames covered twice / once
    cnt = (inv >= 0).sum(1).tolist()
    assert cnt == [2] * 16 + [1] * 4 + [2] * 4 + [1] * 16 + [2] * 4 + [1] * 4
    for f in range(48):
        for j in inv[f].tolist():
            if j >= 0:
                assert windows[j // 24][j % 24] == f
