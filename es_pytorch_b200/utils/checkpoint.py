"""Run-state checkpoints: what the reference's ``Policy.save`` leaves out.

``Policy.save`` / ``Policy.load`` (src/core/policy.py:37-47) pickle the policy with its optimizer moments and obs statistics;
resuming a run bit-for-bit also needs the per-rank RandomState streams (index draws, coins, action noise all come from them,
cached gaussian included), the noise-table seed and the generation counter.  ``save_run_state`` writes them next to the
policy pickle in one file; ``load_run_state`` restores them into fresh objects."""
from __future__ import annotations

import os
import pickle
from typing import Optional, Sequence

import numpy as np


def save_run_state(folder: str, suffix: str, policy, streams: Sequence[np.random.RandomState], table_seed=None,
                   generation: Optional[int] = None, extra: Optional[dict] = None) -> str:
    """-> ``<folder>/run-<suffix>``: {'policy': Policy, 'streams': [RandomState.get_state()], 'table_seed', 'generation', 'extra'}."""
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, f'run-{suffix}')
    state = dict(policy=policy, streams=[rs.get_state() for rs in streams], table_seed=table_seed, generation=generation,
                 extra=dict(extra or {}))
    with open(path, 'wb') as f:
        pickle.dump(state, f)
    return path


def load_run_state(path: str) -> dict:
    """The dict ``save_run_state`` wrote, with ``streams`` rebuilt as RandomState objects at the saved positions."""
    with open(path, 'rb') as f:
        state = pickle.load(f)
    streams = []
    for st in state['streams']:
        rs = np.random.RandomState()
        rs.set_state(st)
        streams.append(rs)
    state['streams'] = streams
    return state
