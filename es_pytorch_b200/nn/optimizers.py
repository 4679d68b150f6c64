"""Optimizers (mirror of src/nn/optimizers.py:7-61): ``Adam``, ``SGD``, ``SimpleES`` with the
reference's constructor signatures and mutable ``lr`` / ``t``.

State (Adam m/v, SGD v) lives in HBM; the update itself is one fused kernel
(``es_adam_step`` & co: /2K, l2 term, moments, theta += step, float32 with one rounding
per reference operation -- numpy 1.18 casting).  ``step(g)`` keeps the reference's
host-array contract for callers that use the optimizer directly.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod

import numpy as np
import torch


class Optimizer(ABC):
    kind = 'abstract'
    _state_names = ()

    def __init__(self, dim: int, lr: float):
        self.lr: float = lr
        self.dim: int = dim
        self.t: int = 0
        self._dev = {}          # name -> float32 device tensor

    # -- device state ------------------------------------------------------------------------
    def _engine(self):
        from ..engine import get_engine
        return get_engine()

    def state(self, name: str) -> torch.Tensor:
        t = self._dev.get(name)
        if t is None:
            eng = self._engine()
            t = self._dev[name] = torch.zeros(self.dim, dtype=torch.float32, device=eng.device)
        return t

    def step(self, globalg):
        """optimizers.py:13-21: increments ``t`` and returns the parameter step for ``globalg``."""
        eng = self._engine()
        g = eng.to_device(np.ascontiguousarray(globalg, dtype=np.float32))
        delta = torch.zeros(self.dim, dtype=torch.float32, device=eng.device)
        # g_total = 0*theta - g/(-1) = g exactly, theta(=0) += step  ->  delta holds the step
        self.apply_fused(eng, delta, g, n_ranked=-1.0, l2coeff=0.0)
        return delta.cpu().numpy()

    def apply_fused(self, eng, theta: torch.Tensor, gsum: torch.Tensor, n_ranked: float, l2coeff: float):
        """theta += step(l2coeff*theta - gsum/n_ranked) in one kernel (es.py:100-101, policy.py:73-74)."""
        self.t += 1
        self._launch(eng, theta, gsum, float(n_ranked), float(l2coeff))

    @abstractmethod
    def _launch(self, eng, theta, gsum, n_ranked, l2coeff):
        pass

    # -- pickling (Policy.save pickles the optimizer, policy.py:43-47) --------------------------------
    # The pickled state is the reference's attribute set: lr, dim, t, hyper-parameters and the moment vectors as float32
    # ndarrays under their reference names (Adam: m, v; SGD: v -- optimizers.py:39,51-52), so checkpoints written by the
    # reference load here with their moments, not with zeros.
    def __getstate__(self):
        d = {k: v for k, v in self.__dict__.items() if k not in ('_dev', '_host_restore')}
        pending = self.__dict__.get('_host_restore') or {}
        for name in self._state_names:
            if name in self._dev:
                d[name] = self._dev[name].cpu().numpy()
            elif name in pending:                                   # restored but not yet used: keep, do not zero
                d[name] = np.array(pending[name], dtype=np.float32)
            else:
                d[name] = np.zeros(self.dim, dtype=np.float32)
        return d

    def __setstate__(self, d):
        d = dict(d)
        host = dict(d.pop('_dev', None) or {})                      # round-1 format of this package
        host.update(d.pop('_host_restore', None) or {})
        for name in self._state_names:
            if name in d:                                           # reference format (and this package's current one)
                host[name] = np.ascontiguousarray(d.pop(name), dtype=np.float32)
        self.__dict__.update(d)
        self._dev = {}
        self._host_restore = host

    def _restore(self):
        host = self.__dict__.pop('_host_restore', None)
        if host:
            eng = self._engine()
            for k, v in host.items():
                self._dev[k] = eng.to_device(v)


class SimpleES(Optimizer):
    kind = 'simple'

    def __init__(self, dim: int, lr: float):
        super().__init__(dim, lr)

    def _launch(self, eng, theta, gsum, n_ranked, l2coeff):
        eng.simple_step(theta, gsum, n_ranked, l2coeff, self.lr)


class SGD(Optimizer):
    kind = 'sgd'
    _state_names = ('v',)

    def __init__(self, dim: int, lr: float, momentum=0.9):
        Optimizer.__init__(self, dim, lr)
        self.momentum = momentum

    @property
    def v(self) -> np.ndarray:
        self._restore()
        return self.state('v').cpu().numpy()

    def _launch(self, eng, theta, gsum, n_ranked, l2coeff):
        self._restore()
        eng.sgd_step(theta, self.state('v'), gsum, n_ranked, l2coeff, self.lr, self.momentum)


class Adam(Optimizer):
    kind = 'adam'
    _state_names = ('m', 'v')

    def __init__(self, dim: int, lr: float, beta1=0.9, beta2=0.999, epsilon=1e-08):
        Optimizer.__init__(self, dim, lr)
        self.beta1 = beta1
        self.beta2 = beta2
        self.epsilon = epsilon

    @property
    def m(self) -> np.ndarray:
        self._restore()
        return self.state('m').cpu().numpy()

    @property
    def v(self) -> np.ndarray:
        self._restore()
        return self.state('v').cpu().numpy()

    def _launch(self, eng, theta, gsum, n_ranked, l2coeff):
        self._restore()
        a = self.lr * math.sqrt(1 - self.beta2 ** self.t) / (1 - self.beta1 ** self.t)    # optimizers.py:56
        eng.adam_step(theta, self.state('m'), self.state('v'), gsum, n_ranked, l2coeff, -a, self.beta1, self.beta2,
                      self.epsilon)
