"""Config / seeding helpers and the gradient reconstruction entry (mirror of src/utils/utils.py)."""
from __future__ import annotations

import argparse
import json
from typing import Optional, Tuple

import numpy as np
import torch


def batch_noise(inds: np.ndarray, nt, policy_len: int, batch_size: int):
    """utils.py:14-26: dense [B, P] copies of table slices, B <= batch_size.  Kept for callers
    that want the rows themselves; the gradient path never materialises them."""
    assert inds.ndim == 1
    rows = []
    for idx in inds:
        rows.append(nt.get(int(idx), policy_len))
        if len(rows) == batch_size:
            yield np.array(rows)
            rows = []
    if rows:
        yield np.array(rows)


def scale_noise_device(engine, fits: torch.Tensor, noise_inds: torch.Tensor, nt, policy_len: int,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum_k fits[k] * table[inds[k] : inds[k]+P] as ONE streaming kernel over the HBM-resident
    table (es_grad_reconstruct) instead of batched gather + sgemv."""
    return engine.grad_reconstruct(nt.device_table(engine), noise_inds, fits, policy_len, out)


def scale_noise(fits: np.ndarray, noise_inds: np.ndarray, nt, policy_len: int, batch_size: int):
    """utils.py:29-39.  ``batch_size`` only bounded the reference's host memory; it is ignored."""
    assert len(fits) == len(noise_inds)
    from ..engine import get_engine
    eng = get_engine()
    w = eng.to_device(np.ascontiguousarray(fits, dtype=np.float32))
    idx = eng.to_device(np.ascontiguousarray(noise_inds).astype(np.int64))
    return scale_noise_device(eng, w, idx, nt, policy_len).cpu().numpy()


def parse_args():
    parser = argparse.ArgumentParser(description='es-pytorch')
    parser.add_argument('config', type=str, help='Config file that will be used')
    return parser.parse_args().config


def load_config(cfg_file: str):
    """JSON -> attribute-access config (utils.py:48-53)."""
    from munch import munchify
    with open(cfg_file) as f:
        return munchify(json.load(f))


def generate_seed(comm) -> int:
    from .. import dist
    return dist.world().broadcast_object(int(np.random.randint(0, 1000000)), 0)


def seed(comm, seed: list, env=None) -> Tuple[np.random.RandomState, int, int]:
    """Per-rank RandomState + torch seed shared by all ranks (utils.py:61-76)."""
    from .. import dist
    if seed is not None and hasattr(seed, '__len__') and len(seed) == comm.size:
        my_seed = seed[comm.rank]
        rs = np.random.RandomState(my_seed)
    else:
        my_seed = int(np.random.SeedSequence().generate_state(1)[0] % (2 ** 31))
        rs = np.random.RandomState(my_seed)
    global_seed = dist.world().broadcast_object(my_seed, 0) if comm.size > 1 else my_seed
    torch.random.manual_seed(global_seed)   # identical initial params on every rank
    if env is not None:
        env.seed(my_seed)
        env.action_space.seed(my_seed)
        env.observation_space.seed(my_seed)
    return rs, my_seed, global_seed
