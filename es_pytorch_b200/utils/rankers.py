"""Fitness rank transforms (mirror of src/utils/rankers.py).

Every ranker of the reference is a thin description (shaping kind, objective blend, elite
fraction) of one device call, ``es_rank_transform``: integer-exact ranks with a stable tie
order, and one IEEE operation per reference operation in the reference's dtype (float32
for the rank-based shapings, float64 for ``MaxNormalizedRanker``).  ``rank`` accepts the
reference's host arrays (uploaded, 8 bytes per fitness) or arrays that shadow device
tensors (``devcache``); ``rank_device`` is the device-resident variant used by
``DeviceGeneration``.

Deviations, both forced by behaviour the reference leaves unspecified:
  * ties are broken by position (numpy's default argsort is unstable, rankers.py:16);
  * ``EliteRanker`` returns its elite in ascending rank order (``np.argpartition``'s order
    is unspecified, rankers.py:95); the elite *set* is the same whenever it is well defined.
"""
from __future__ import annotations

from abc import ABC
from typing import Optional

import numpy as np
import torch

from .._lib import (ES_RANK_CENTERED, ES_RANK_DOUBLE_POSITIVE, ES_RANK_MAX_NORMALIZED, ES_RANK_SEMI_CENTERED)


def _as_2d(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(len(a), -1)


class Ranker(ABC):
    """Ranks all fitnesses obtained in a generation (rankers.py:20-50)."""

    kind: Optional[int] = None          # ES_RANK_* shaping of a plain (single objective) ranker
    squeeze = True                      # CenteredRanker / MaxNormalizedRanker np.squeeze their result; SemiCentered does not

    def __init__(self):
        self.fits_pos: Optional[np.ndarray] = None
        self.fits_neg: Optional[np.ndarray] = None
        self.noise_inds: Optional[np.ndarray] = None
        self.ranked_fits: Optional[np.ndarray] = None
        self.n_fits_ranked: int = 0
        self.ranked_fits_dev: Optional[torch.Tensor] = None

    fits = property(lambda self: np.concatenate((self.fits_pos, self.fits_neg)))

    # -- description of the device call ------------------------------------------------------
    def _spec(self, n_obj: int, n_fits: int):
        """(kind, w0, w1, elite_n) for ``n_obj`` objective columns and ``n_fits`` = 2K fitnesses."""
        if self.kind is None:
            raise NotImplementedError(f'{type(self).__name__} does not describe a device rank transform')
        if n_obj != 1:
            raise ValueError(f'{type(self).__name__} ranks a single objective; wrap it in MultiObjectiveRanker for two')
        return self.kind, 1.0, 0.0, 0

    def _pre_rank(self, fits_pos, fits_neg, noise_inds):
        self.fits_pos, self.fits_neg, self.noise_inds = fits_pos, fits_neg, noise_inds

    def rank_device(self, engine, fpos: torch.Tensor, fneg: torch.Tensor, k_begin: int = 0,
                    k_count: Optional[int] = None) -> torch.Tensor:
        """Device-resident variant: float64 [K, n_obj] tensors in, float32 weights of the pairs
        [k_begin, k_begin+k_count) out (one weight per pair, also for EliteRanker: the sum of the pair's elite
        values); ranks are global over all K pairs.  Sets ``n_fits_ranked`` like the reference."""
        n_obj = 1 if fpos.dim() == 1 else fpos.shape[1]
        K = fpos.shape[0]
        kind, w0, w1, elite_n = self._spec(n_obj, 2 * K)
        self.n_fits_ranked = elite_n if elite_n else 2 * K           # ranked_fits.size, rankers.py:43,102
        out = engine.rank_transform(fpos, fneg, kind, w0, w1, elite_n, k_begin, k_count)
        self.ranked_fits_dev = out['weights']
        return self.ranked_fits_dev

    def rank(self, fits_pos: np.ndarray, fits_neg: np.ndarray, noise_inds: np.ndarray) -> np.ndarray:
        from .. import devcache
        from ..engine import get_engine
        eng = get_engine()
        self._pre_rank(fits_pos, fits_neg, noise_inds)
        fp, fn = devcache.lookup(fits_pos), devcache.lookup(fits_neg)
        if fp is None or fn is None:
            fp = eng.to_device(_as_2d(fits_pos), torch.float64)
            fn = eng.to_device(_as_2d(fits_neg), torch.float64)
        n_obj = 1 if fp.dim() == 1 else fp.shape[1]
        K = fp.shape[0]
        kind, w0, w1, elite_n = self._spec(n_obj, 2 * K)
        f64 = kind == ES_RANK_MAX_NORMALIZED
        if elite_n:
            return self._rank_elite(eng, fp, fn, kind, elite_n, f64)
        out = eng.rank_transform(fp, fn, kind, w0, w1, 0, want64=f64)
        self.n_fits_ranked = 2 * K
        w = out['weights']
        self.ranked_fits_dev = w
        h = eng.download_async(out['weights64'] if f64 else w, ('ranked', id(self)))
        eng.sync()
        res = h.numpy().copy()
        if not self._squeezes() and np.ndim(fits_pos) == 2:
            res = res.reshape(-1, 1)                                 # SemiCenteredRanker keeps [2K, 1] (rankers.py:80-83)
        self.ranked_fits = devcache.attach(res, w, lambda r=self, t=w: r.ranked_fits_dev is t)
        return self.ranked_fits

    def _squeezes(self) -> bool:
        return self.squeeze

    def _rank_elite(self, eng, fp, fn, kind, elite_n, f64):
        from .. import devcache
        K = fp.shape[0]
        inds_dev = devcache.lookup(self.noise_inds)
        if inds_dev is None or inds_dev.dtype != torch.int64 or inds_dev.numel() != K:
            inds_dev = eng.to_device(np.ascontiguousarray(self.noise_inds).astype(np.int64))
        out = eng.rank_transform(fp, fn, kind, 1.0, 0.0, elite_n, noise_idx=inds_dev, want_elite=True)
        self.n_fits_ranked = elite_n
        vals = out['elite_vals'] if f64 else out['elite_vals'].to(torch.float32)   # exact: the values are float32
        self.ranked_fits_dev = vals if not f64 else vals.to(torch.float32)
        hv = eng.download_async(vals, ('elite_vals', id(self)))
        hf = eng.download_async(out['elite_fit'], ('elite_fit', id(self)))
        eng.sync()
        fit = hf.numpy().astype(np.int64)
        dev_w, dev_i = self.ranked_fits_dev, out['elite_idx']
        # setting the noise inds to only be the inds of the elite (rankers.py:96-97)
        sel = np.asarray(self.noise_inds)[fit % len(self.noise_inds)]
        self.noise_inds = devcache.attach(sel, dev_i, lambda r=self, t=dev_w: r.ranked_fits_dev is t)
        self.ranked_fits = devcache.attach(hv.numpy().copy(), dev_w, lambda r=self, t=dev_w: r.ranked_fits_dev is t)
        return self.ranked_fits


class CenteredRanker(Ranker):
    """rank -> float32(rank)/(n-1) - 0.5 -> pos minus neg (rankers.py:53-58,42-44)."""
    kind = ES_RANK_CENTERED


class DoublePositiveCenteredRanker(CenteredRanker):
    """Centered ranks with the positive half doubled (rankers.py:61-65)."""
    kind = ES_RANK_DOUBLE_POSITIVE


class MaxNormalizedRanker(Ranker):
    """Raw fitnesses shifted by their minimum, divided by the maximum and stretched to [-1, 1], float64
    (rankers.py:68-75)."""
    kind = ES_RANK_MAX_NORMALIZED


class SemiCenteredRanker(Ranker):
    """((1/s) * (rank + 0.29 s)^2) / s - 0.5 in float32 (rankers.py:78-83)."""
    kind = ES_RANK_SEMI_CENTERED
    squeeze = False


def _plain(ranker: Ranker, who: str) -> Ranker:
    if not isinstance(ranker, Ranker) or ranker.kind is None:
        raise NotImplementedError(f'{who} needs a plain shaping ranker (Centered / DoublePositiveCentered / '
                                  f'SemiCentered / MaxNormalized), got {type(ranker).__name__}')
    return ranker


class EliteRanker(Ranker):
    """Keeps only the ``elite_percent`` best shaped fitnesses, unsubtracted, each with the noise index of its pair
    (rankers.py:86-103)."""

    def __init__(self, ranker: Ranker, elite_percent: float):
        super().__init__()
        assert 0 <= elite_percent <= 1
        self.ranker = _plain(ranker, 'EliteRanker')
        self.elite_percent = elite_percent

    def _spec(self, n_obj: int, n_fits: int):
        kind, w0, w1, _ = self.ranker._spec(n_obj, n_fits)
        return kind, w0, w1, max(1, int(n_fits * self.elite_percent))      # rankers.py:94


class MultiObjectiveRanker(Ranker):
    """Two objective columns shaped independently and blended w*r0 + (1-w)*r1 (rankers.py:106-120)."""

    def __init__(self, ranker: Ranker, w: float):
        assert 0. <= w <= 1.
        super().__init__()
        self.ranker = _plain(ranker, 'MultiObjectiveRanker')
        self.w = w

    def _spec(self, n_obj: int, n_fits: int):
        assert n_obj == 2  # this only works for 2 objectives (rankers.py:114)
        return self.ranker.kind, self.w, 1 - self.w, 0

    def _squeezes(self) -> bool:
        return True        # the blend of two columns is 1-D whatever the inner ranker returns
