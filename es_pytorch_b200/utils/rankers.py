"""Fitness rank transforms (mirror of src/utils/rankers.py).

On the hot path: ``CenteredRanker`` and ``MultiObjectiveRanker`` -- ranks and the float32
affine map are computed by ``es_centered_rank`` (integer-exact ranks, stable tie order,
one IEEE float32 operation per reference operation).  ``rank`` accepts the reference's
host arrays (uploaded, 8 bytes per fitness) or device tensors.
The remaining rankers of the reference (DoublePositive / MaxNormalized / SemiCentered /
Elite) are outside the hot-path scope and are not provided yet.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch


def _as_2d(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(len(a), -1)


class Ranker(ABC):
    """Ranks all fitnesses obtained in a generation (rankers.py:20-50)."""

    def __init__(self):
        self.fits_pos: Optional[np.ndarray] = None
        self.fits_neg: Optional[np.ndarray] = None
        self.noise_inds: Optional[np.ndarray] = None
        self.ranked_fits: Optional[np.ndarray] = None
        self.n_fits_ranked: int = 0
        self.ranked_fits_dev: Optional[torch.Tensor] = None

    fits = property(lambda self: np.concatenate((self.fits_pos, self.fits_neg)))

    @abstractmethod
    def _blend(self, n_obj: int):
        """(w0, w1) blend of the per-objective centered ranks."""

    def _pre_rank(self, fits_pos, fits_neg, noise_inds):
        self.fits_pos, self.fits_neg, self.noise_inds = fits_pos, fits_neg, noise_inds

    def rank_device(self, engine, fpos: torch.Tensor, fneg: torch.Tensor, k_begin: int = 0,
                    k_count: Optional[int] = None) -> torch.Tensor:
        """Device-resident variant: float64 [K, n_obj] tensors in, float32 weights of the pairs
        [k_begin, k_begin+k_count) out; ranks are global over all K pairs."""
        n_obj = 1 if fpos.dim() == 1 else fpos.shape[1]
        w0, w1 = self._blend(n_obj)
        self.n_fits_ranked = 2 * fpos.shape[0]                      # ranked_fits.size, rankers.py:43
        self.ranked_fits_dev = engine.centered_rank(fpos, fneg, w0, w1, k_begin, k_count)
        return self.ranked_fits_dev

    def rank(self, fits_pos: np.ndarray, fits_neg: np.ndarray, noise_inds: np.ndarray) -> np.ndarray:
        from ..engine import get_engine
        eng = get_engine()
        self._pre_rank(fits_pos, fits_neg, noise_inds)
        from .. import devcache
        fp, fn = devcache.lookup(fits_pos), devcache.lookup(fits_neg)
        if fp is None or fn is None:
            fp = eng.to_device(_as_2d(fits_pos), torch.float64)
            fn = eng.to_device(_as_2d(fits_neg), torch.float64)
        w = self.rank_device(eng, fp, fn)
        h = eng.download_async(w, ('ranked', id(self)))
        eng.sync()
        self.ranked_fits = devcache.attach(h.numpy().copy(), w, lambda r=self, t=w: r.ranked_fits_dev is t)
        return self.ranked_fits


class CenteredRanker(Ranker):
    """rank -> float32(rank)/(n-1) - 0.5 -> pos minus neg (rankers.py:53-58,42-44)."""

    def _blend(self, n_obj: int):
        if n_obj != 1:
            raise ValueError('CenteredRanker ranks a single objective; wrap it in MultiObjectiveRanker for two')
        return 1.0, 0.0


class MultiObjectiveRanker(Ranker):
    """Two objective columns ranked independently and blended w*r0 + (1-w)*r1 (rankers.py:106-120)."""

    def __init__(self, ranker: Ranker, w: float):
        assert 0. <= w <= 1.
        super().__init__()
        if not isinstance(ranker, CenteredRanker):
            raise NotImplementedError('only MultiObjectiveRanker(CenteredRanker(), w) is on the device path')
        self.ranker = ranker
        self.w = w

    def _blend(self, n_obj: int):
        assert n_obj == 2  # this only works for 2 objectives (rankers.py:114)
        return self.w, 1 - self.w


class _NotOnDevicePath(Ranker):
    """Rankers of the reference that are outside the hot-path scope.
    The names exist so that scripts importing them load; using one raises instead of silently
    running a host implementation."""

    def _blend(self, n_obj: int):
        raise NotImplementedError(f'{type(self).__name__} is not implemented on the device path yet '
                                  f'(reference: src/utils/rankers.py:61-103)')


class DoublePositiveCenteredRanker(_NotOnDevicePath):
    pass


class MaxNormalizedRanker(_NotOnDevicePath):
    pass


class SemiCenteredRanker(_NotOnDevicePath):
    pass


class EliteRanker(_NotOnDevicePath):
    def __init__(self, ranker: Ranker, elite_percent: float):
        super().__init__()
        assert 0 <= elite_percent <= 1
        self.ranker = ranker
        self.elite_percent = elite_percent
