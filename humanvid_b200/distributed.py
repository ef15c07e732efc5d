"""Multi-GPU entry points of the denoising path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests) for the plumbing.

The reference's inference path has no collective at all (it is single-GPU Python).  Two partitions exist in the path:

* **clips** (BASELINE config 4): independent clips, one per rank, weights replicated, no communication during the 25 steps,
  then ONE all-gather of the final latents (``gather_clip_latents``) before the VAE decode.
* **(context window x CFG half) units of ONE clip** (config 5: 48 frames = 3 windows, x 2 CFG halves = 6 units per timestep;
  config 2/3: 1 window x 2 halves = the 2-GPU CFG split).  Every unit is an independent UNet forward at batch 1
  (``HV_FLAG_UNCOND_ONLY`` / ``HV_FLAG_COND_ONLY``); the unit predictions (<= 7.1 MB per step at 576x1024) are exchanged with one
  all-gather per timestep and every rank then runs the tiny accumulate / CFG / DDIM kernel redundantly, so the latents stay
  replicated without a broadcast (pipeline_pose2vid_long.py:494-563, context.py:15-42).  With more ranks than units the extra
  ranks idle (6 units on 8 GPUs: ideal speed-up 6x); splitting one window further needs an all-to-all around each of the 21
  temporal modules and is not built.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .pipeline import Pose2VideoPipeline

Unit = Tuple[int, int]   # (window index, CFG half: 0 = unconditional, 1 = conditional)


def unit_list(n_windows: int, cfg_on: bool) -> List[Unit]:
    return [(w, h) for w in range(n_windows) for h in ((0, 1) if cfg_on else (0,))]


UNIT_COST = (1.0, 1.12)   # relative device time of an unconditional / conditional unit: the conditional half also attends over the reference banks


def assign_units(units: Sequence[Unit], world: int) -> List[List[Unit]]:
    """Longest-processing-time-first: the heavier conditional units are placed first, every unit goes to the least loaded rank (ties: lowest
    rank).  6 units: 2 ranks -> {c0, c2, u2} + {c1, u0, u1}, 4 ranks -> 2 + 1 + 1 + 2, 8 ranks -> six busy.  Round-robin would put all three
    conditional units of a 2-rank split on one rank (measured: 1.73x instead of the ideal 2x at config 5).  The result is deterministic and the
    same on every rank; which rank computes a unit does not change its value (tests/test_multigpu_gpu.py: bitwise)."""
    out: List[List[Unit]] = [[] for _ in range(world)]
    load = [0.0] * world
    for u in sorted(units, key=lambda u: (-UNIT_COST[u[1]], u[0])):
        r = min(range(world), key=lambda i: (load[i], i))
        out[r].append(u)
        load[r] += UNIT_COST[u[1]]
    for a in out:
        a.sort()
    return out


class UnitExchange:
    """One all-gather per timestep: every rank contributes the predictions of its units (padded to the per-rank maximum)
    and receives all of them.  ``__call__(mine)`` -> {unit: tensor}; the tensors are views of one gathered buffer."""

    def __init__(self, units: Sequence[Unit], world: Optional[int] = None, rank: Optional[int] = None, group=None):
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.group = group
        self.units = list(units)
        self.assignment = assign_units(self.units, self.world)
        self.slots = max(1, max(len(a) for a in self.assignment))
        self.where: Dict[Unit, Tuple[int, int]] = {u: (r, s) for r, a in enumerate(self.assignment) for s, u in enumerate(a)}
        self._send = None
        self._recv = None

    @property
    def my_units(self) -> List[Unit]:
        return self.assignment[self.rank]

    def __call__(self, mine: Sequence[torch.Tensor]) -> Dict[Unit, torch.Tensor]:
        if len(mine) != len(self.my_units):
            raise ValueError(f"rank {self.rank} owns {len(self.my_units)} units, got {len(mine)} predictions")
        shape = self._shape = tuple(mine[0].shape) if mine else self._shape
        ref = mine[0] if mine else self._recv
        if self._send is None or tuple(self._send.shape[1:]) != shape:
            self._send = torch.zeros((self.slots, *shape), device=ref.device, dtype=ref.dtype)
            self._recv = torch.empty((self.world * self.slots, *shape), device=ref.device, dtype=ref.dtype)
        for s, t in enumerate(mine):
            self._send[s].copy_(t)
        if self.world > 1:
            dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        else:
            self._recv.copy_(self._send)
        return {u: self._recv[r * self.slots + s] for u, (r, s) in self.where.items()}

    def prime(self, shape, device, dtype):
        """Ranks that own no unit (more ranks than units) still take part in the collective: give them the buffer shape."""
        self._shape = tuple(shape)
        self._send = torch.zeros((self.slots, *shape), device=device, dtype=dtype)
        self._recv = torch.empty((self.world * self.slots, *shape), device=device, dtype=dtype)


def gather_clip_latents(latents: torch.Tensor, group=None) -> torch.Tensor:
    """BASELINE config 4: every rank denoised its own clip (1, 4, F, h, w); reassemble (world, 4, F, h, w) on every rank with
    the path's single collective (1.33 MB per rank at 24x96x72 fp16)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return latents.clone()
    out = torch.empty((world * latents.shape[0], *latents.shape[1:]), device=latents.device, dtype=latents.dtype)
    dist.all_gather_into_tensor(out, latents.contiguous(), group=group)
    return out


class ShardedPose2VideoPipeline(Pose2VideoPipeline):
    """``Pose2VideoPipeline`` for ONE clip over several GPUs: identical call on every rank (same seed -> same initial latents, same
    reference banks), the per-timestep UNet work split into (window x CFG-half) units, latents replicated.  Returns the same video
    on every rank."""

    def __init__(self, *a, process_group=None, **k):
        super().__init__(*a, **k)
        self.process_group = process_group

    def _loop_kwargs(self, context_queue, cfg_on):
        if not dist.is_initialized() or dist.get_world_size(self.process_group) == 1:
            return {}
        units = unit_list(len(context_queue), cfg_on)
        ex = UnitExchange(units, group=self.process_group)
        return {"exchange": ex, "my_units": ex.my_units, "all_units": units}
