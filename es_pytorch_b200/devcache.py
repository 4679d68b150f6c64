"""Device shadows of the numpy arrays the reference API hands around.

``es.test_params`` must return ndarrays (reference contract), and scripts pass them straight back into
``Ranker.rank`` / ``es.approx_grad``.  To avoid re-uploading what the GPU produced a moment ago, an array returned by
this package can carry a *device shadow*: the array is made read-only (so its contents cannot diverge from the shadow)
and registered here by identity together with a validity token; a lookup only succeeds for the very same object while
the token is still current (the shadow buffers are reused by the next generation).
"""
from __future__ import annotations

import weakref
from typing import Callable, Optional

import numpy as np
import torch

_REG = {}


def attach(arr: np.ndarray, dev: torch.Tensor, still_valid: Callable[[], bool]) -> np.ndarray:
    arr.flags.writeable = False
    key = id(arr)
    _REG[key] = (weakref.ref(arr, lambda _r, k=key: _REG.pop(k, None)), dev, still_valid)
    return arr


def lookup(arr) -> Optional[torch.Tensor]:
    e = _REG.get(id(arr))
    if e is None or e[0]() is not arr or not e[2]():
        return None
    return e[1]
