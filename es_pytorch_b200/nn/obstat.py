"""Running observation statistics (mirror of src/nn/obstat.py:13-43).

A small host-side float64 record (sum, sumsq, count) -- the API object scripts pass around.
The per-generation accumulation over saved rollouts runs on the device
(``es_obstat_accumulate_coins``); this class only holds / merges the results.
"""
from __future__ import annotations

import numpy as np


class ObStat:
    def __init__(self, shape, eps):
        self.sum = np.zeros(shape, dtype=np.float64)
        self.sumsq = np.full(shape, eps, dtype=np.float64)
        self.count = eps

    def inc(self, s, ssq, c):
        """obstat.py:19-22."""
        self.sum += np.asarray(s).astype(np.float64)
        self.sumsq += np.asarray(ssq).astype(np.float64)
        self.count += c

    def __iadd__(self, other: 'ObStat'):
        self.inc(other.sum, other.sumsq, other.count)
        return self

    def __repr__(self):
        return f'sum:{self.sum} sumsq:{self.sumsq} count:{self.count}'

    @property
    def mean(self):
        return self.sum / self.count

    @property
    def std(self):
        """obstat.py:35-37: sqrt(max(E[x^2] - E[x]^2, 1e-2))."""
        return np.sqrt(np.maximum(self.sumsq / self.count - np.square(self.mean), 1e-2))

    def mpi_inc(self, comm):
        """Sum the record over all ranks (obstat.py:39-43; a pickled custom-op allreduce in
        the reference, a float64 tensor allreduce here)."""
        if getattr(comm, 'size', 1) == 1:
            return
        import torch
        from .. import dist
        packed = np.concatenate([self.sum.ravel(), self.sumsq.ravel(), [float(self.count)]])
        t = torch.from_numpy(packed)
        if torch.cuda.is_available() and torch.distributed.get_backend() == 'nccl':
            t = t.cuda()
        dist.world().allreduce_sum(t)
        packed = t.cpu().numpy()
        n = self.sum.size
        self.sum = packed[:n].reshape(self.sum.shape)
        self.sumsq = packed[n:2 * n].reshape(self.sumsq.shape)
        self.count = float(packed[-1])
