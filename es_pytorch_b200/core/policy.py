"""Policy = flat parameter vector <-> module (mirror of src/core/policy.py:19-74).

``flat_params`` (float32 ndarray, mutated in place -- the reference contract) is mirrored
by ``theta_dev`` in HBM; perturbation arithmetic (theta + sigma*eps, policy.py:64) and the
optimizer step run on the device.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from ..nn.nn import BaseNet
from ..nn.obstat import ObStat
from ..nn.optimizers import Optimizer


def init_normal(m):
    if type(m) == torch.nn.Linear:
        torch.nn.init.kaiming_normal_(m.weight)


class Policy:
    _theta_dev = None          # class-level default: a Policy pickled by the reference has no such attribute

    def __init__(self, module: BaseNet, noise_std: float, optim: Optimizer):
        module.apply(init_normal)
        self._module: BaseNet = module
        self.std = noise_std
        self.flat_params: np.ndarray = Policy.get_flat(module)
        self.obstat: ObStat = ObStat(module._obmean.shape, 1e-2)
        self.optim = optim
        self._theta_dev = None

    def __len__(self):
        return len(self.flat_params)

    @staticmethod
    def get_flat(module: torch.nn.Module) -> np.ndarray:
        """state_dict order, each tensor flattened row-major (policy.py:33-35)."""
        return torch.cat([t.detach().flatten().cpu() for t in module.state_dict().values()]).numpy().copy()

    # -- device mirror ----------------------------------------------------------------------------
    def theta_dev(self, engine, refresh: bool = True) -> torch.Tensor:
        """theta in HBM; ``refresh`` re-uploads ``flat_params`` (117 kB) so host-side edits made
        by scripts between generations are honoured."""
        if self._theta_dev is None or self._theta_dev.device != engine.device:
            self._theta_dev = engine.to_device(self.flat_params, torch.float32)
        elif refresh:
            engine.upload_async(self._theta_dev, self.flat_params, ('theta', id(self)))
        return self._theta_dev

    def sync_host(self):
        """Copy theta back into ``flat_params`` in place (keeps aliases held by scripts valid)."""
        if self._theta_dev is not None:
            from ..engine import get_engine
            eng = get_engine(self._theta_dev.device.index)
            h = eng.download_async(self._theta_dev, ('theta', id(self)))
            eng.sync()
            self.flat_params[...] = h.numpy()

    # -- checkpointing (policy.py:37-47) -------------------------------------------------------------
    @staticmethod
    def load(file: str) -> 'Policy':
        with open(file, 'rb') as f:
            policy: Policy = pickle.load(f)
        policy.set_nn_params(policy.flat_params)
        return policy

    def save(self, folder: str, suffix: str):
        if not os.path.exists(folder):
            os.makedirs(folder)
        with open(os.path.join(folder, f'policy-{suffix}'), 'wb') as f:
            pickle.dump(self, f)

    def __getstate__(self):
        # exactly the reference's attribute set (policy.py:23-29): _module, std, flat_params, obstat, optim
        d = dict(self.__dict__)
        d.pop('_theta_dev', None)
        d.pop('_sd_views', None)
        return d

    def __setstate__(self, d):
        """Accepts this package's pickles and the reference's (``src.core.policy.Policy`` resolved through the compat
        shims): device mirrors are rebuilt lazily."""
        self.__dict__.update(d)
        self._theta_dev = None
        self.flat_params = np.ascontiguousarray(self.flat_params, dtype=np.float32)

    # -- reference API -----------------------------------------------------------------------------------
    def set_nn_params(self, params) -> torch.nn.Module:
        """Scatter a flat vector into the module tensors (policy.py:49-59)."""
        flat = params if isinstance(params, torch.Tensor) else torch.from_numpy(np.asarray(params))
        with torch.no_grad():
            at = 0
            # state_dict order = flat order (policy.py:33-35); copied in place: the effect of the reference's
            # load_state_dict without rebuilding the dict machinery every evaluation.  The tensor list is cached (it
            # aliases the module's parameters and buffers) and rebuilt if the module was swapped or restructured.
            ptrs = tuple(p.data_ptr() for p in self._module.parameters())
            cache = getattr(self, '_sd_views', None)
            if cache is None or cache[0] is not self._module or cache[2] != ptrs:
                cache = self._sd_views = (self._module, list(self._module.state_dict().values()), ptrs)
            for w in cache[1]:
                n = w.numel()
                w.copy_(flat[at:at + n].reshape(w.shape))
                at += n
        return self._module

    def pheno(self, noise: np.ndarray = None) -> torch.nn.Module:
        """Module carrying theta + std*noise (policy.py:61-67).  The arithmetic runs on the
        device (es_perturb): the caller's noise vector is uploaded as a one-slice table."""
        from ..engine import get_engine
        eng = get_engine()
        theta = self.theta_dev(eng)
        if noise is None or not np.any(noise):
            params = theta                                   # es.py:48 passes float64 zeros: theta unchanged
        else:
            n32 = np.ascontiguousarray(noise, dtype=np.float32)
            tbl = torch.zeros(len(self) + 1, dtype=torch.float32, device=eng.device)
            tbl[:len(self)].copy_(torch.from_numpy(n32), non_blocking=True)
            idx0 = torch.zeros(1, dtype=torch.int64, device=eng.device)
            params, _ = eng.perturb(theta, tbl, idx0, self.std, want_neg=False)
            params = params.view(-1)
        self.set_nn_params(params.cpu())
        return self._module

    def update_obstat(self, obstat: ObStat):
        self.obstat += obstat
        self._module.set_ob_mean_std(self.obstat.mean, self.obstat.std)

    def optim_step(self, global_g):
        self.flat_params += self.optim.step(global_g)
