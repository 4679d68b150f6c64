"""Novelty helpers (mirror of src/utils/novelty.py:9-18).  The batched per-rollout novelty of
the NSRA path is ``es_novelty`` on the device; these host functions serve the per-generation
archive bookkeeping that scripts do with a handful of 2-D points."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch


def update_archive(comm, behaviour: Sequence[float], archive: Optional[np.ndarray]) -> np.ndarray:
    """Append rank 0's behaviour (novelty.py:9-13)."""
    from .. import dist
    if getattr(comm, 'size', 1) > 1:
        behaviour = dist.world().broadcast_object(behaviour, 0)
    if archive is None:
        return np.array([behaviour])
    return np.concatenate((archive, [behaviour]))


def novelty(behaviour: np.ndarray, archive: np.ndarray, n: int) -> float:
    """Mean of the n smallest euclidean distances to the archive (novelty.py:16-18)."""
    from ..engine import get_engine
    eng = get_engine()
    b = np.zeros((1, 3), dtype=np.float32)
    bh = np.asarray(behaviour, dtype=np.float64).reshape(-1)
    a = np.asarray(archive, dtype=np.float64)
    if bh.size != 2 or a.shape[1] != 2 or np.any(bh != bh.astype(np.float32)):
        raise NotImplementedError('device novelty handles float32-representable 2-D behaviours (positions[-3:-1])')
    b[0, :2] = bh
    out = torch.zeros(1, dtype=torch.float64, device=eng.device)
    eng.novelty(eng.to_device(b), eng.to_device(a, torch.float64), int(n), out, 1)
    return float(out.item())
