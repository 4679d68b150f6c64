"""Process-group plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo
in CPU tests) replaces the reference's mpi4py communicator.

The reference exchanges per generation:
  * ``comm.Alltoall`` used as an allgather of ``[f+..., f-..., idx]`` rows  (src/core/es.py:84-95)
  * ``comm.allreduce`` of step counts and ObStat                          (es.py:78-79, obstat.py:39-43)
and every rank then recomputes the full gradient.  Here each rank reconstructs only its
own shard's partial sum and the ranks do ONE allreduce of the float32[P] partial gradient
(``allreduce_sum``), plus one allgather of the fitness rows so that ranks are global.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as td


class Comm:
    """Minimal communicator with the attributes the reference's code reads from
    ``MPI.COMM_WORLD`` (``rank``, ``size``) plus the tensor collectives this package needs."""

    def __init__(self, group=None):
        self.group = group

    @property
    def active(self) -> bool:
        return td.is_available() and td.is_initialized()

    @property
    def rank(self) -> int:
        return td.get_rank(self.group) if self.active else 0

    @property
    def size(self) -> int:
        return td.get_world_size(self.group) if self.active else 1

    # -- tensor collectives (in place / into preallocated outputs; stream-ordered on NCCL) --
    def allgather_into(self, out: torch.Tensor, local: torch.Tensor) -> torch.Tensor:
        if self.size == 1:
            out.view(-1)[:local.numel()].copy_(local.view(-1))
            return out
        # flat views: rank r's block lands at out[r] for any [size, *local.shape] output (NCCL and gloo alike)
        td.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=self.group)
        return out

    def allreduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            td.all_reduce(t, op=td.ReduceOp.SUM, group=self.group)
        return t

    def barrier(self):
        if self.size > 1:
            td.barrier(group=self.group)

    def broadcast_object(self, obj, src: int = 0):
        if self.size == 1:
            return obj
        box = [obj]
        td.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def allgather_object(self, obj) -> List:
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        td.all_gather_object(out, obj, group=self.group)
        return out


_WORLD: Optional[Comm] = None


def world() -> Comm:
    global _WORLD
    if _WORLD is None:
        _WORLD = Comm()
    return _WORLD


def init_from_env(backend: Optional[str] = None) -> Comm:
    """Initialise the default process group from torchrun's environment (RANK, WORLD_SIZE,
    MASTER_ADDR, MASTER_PORT, LOCAL_RANK).  No-op for a single process."""
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws > 1 and not (td.is_available() and td.is_initialized()):
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        td.init_process_group(backend=backend)
    return world()


def shard_bounds(total: int, size: int, rank: int):
    """Contiguous shard [begin, end) of ``total`` units for ``rank`` of ``size`` (units are
    antithetic pairs; the reference asserts divisibility, src/core/es.py:38)."""
    if total % size != 0:
        raise ValueError(f'{total} pairs do not divide over {size} ranks (es.py:38 asserts divisibility)')
    per = total // size
    return rank * per, (rank + 1) * per
