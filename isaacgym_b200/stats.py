"""Scalar episode statistics (what the reference prints: mean reward / mean progress every
`log_every` steps, TILT:763-766, and the five ADOF counter sums, ADOF:1164-1168).

The step kernel accumulates warp-reduced partial sums into PPK_STATS_SLOTS x PPK_NUM_STATS doubles;
`reduce()` folds them on the device.  Across GPUs the env batch is sharded and the only collective
of the whole path is one SUM all-reduce of these 8 doubles (NCCL over NVLink), issued on a side
stream so it never sits on the step's critical path.  Nothing here synchronises with the host
until `means()` is asked for.
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _native as N


class EpisodeStats:
    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.slots = torch.zeros(N.PPK_STATS_SLOTS, N.PPK_NUM_STATS, dtype=torch.float64, device=self.device)
        self.local = torch.zeros(N.PPK_NUM_STATS, dtype=torch.float64, device=self.device)
        self.total = torch.zeros(N.PPK_NUM_STATS, dtype=torch.float64, device=self.device)
        self.group = group
        self._side: Optional[torch.cuda.Stream] = None
        self._work = None

    def reduce(self, lib, stream_ptr: int):
        """slots -> local[8] (and zero the slots) on the step's stream, then all-reduce."""
        N.check(lib.ppk_stats_reduce(self.slots.data_ptr(), self.local.data_ptr(), stream_ptr), "ppk_stats_reduce")
        self.all_reduce()

    def all_reduce(self):
        """SUM over ranks of the 8-double vector; a no-op for a single process."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            self.total.copy_(self.local, non_blocking=True)
            return
        if self.device.type == "cuda":
            if self._side is None:
                self._side = torch.cuda.Stream(self.device)
            self._side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._side):
                self.total.copy_(self.local, non_blocking=True)
                self._work = dist.all_reduce(self.total, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            self.total.copy_(self.local)
            dist.all_reduce(self.total, op=dist.ReduceOp.SUM, group=self.group)

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._side)

    def means(self, global_num_envs: int) -> Dict[str, float]:
        """Host read (the only sync): per-env means over the global batch."""
        self.wait()
        v = self.total.cpu().tolist()
        return {name: x / global_num_envs for name, x in zip(N.STAT_NAMES, v)}


def shard_range(num_envs: int, rank: int, world: int):
    """Contiguous env block of `rank`: [lo, hi)."""
    base, rem = divmod(num_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
