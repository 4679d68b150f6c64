"""Timed CPU baseline: the reference's generation step on the host cores.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/es_oracle.py's header).  The reference itself
cannot run on the box (mpi4py / gym / munch / mlflow / mpirun absent, numpy 2.x removed
``np.float``, src/core/es.py:89), so this times its line-by-line restatement:

  * ``mpirun -np R`` is emulated by R = len(os.sched_getaffinity(0)) worker processes
    (one per core, ``torch.set_num_threads(1)`` each); the noise table is shared
    copy-on-write through ``fork`` (stands in for the MPI shared window,
    src/core/noisetable.py:13-24);
  * every worker runs the reference loop for a bounded number of antithetic pairs:
    ``nt.sample`` -> ``pheno`` (+ ``load_state_dict``) -> per-step batch-1 torch forward in a
    python loop -> ``sum(rews)`` (src/core/es.py:67-74, src/gym/gym_runner.py:50-54);
    the per-pair time is extrapolated linearly to K/R pairs per rank;
  * rank + scale_noise(batch 500) + Adam run IN FULL at the real K in every worker at once
    (the reference executes them redundantly on every rank, es.py:98-101), so their wall
    time includes R-way memory contention.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_G = {}


def _reference_rollout_pair(table, theta, sigma, module, dims, env, idx, P):
    """One antithetic pair exactly as the reference does it (module load + python step loop)."""
    import torch
    from oracle import es_oracle as orc
    noise = table[idx:idx + P]
    out = []
    for sign in (1.0, -1.0):
        params = orc.pheno_params(theta, sigma, noise if sign > 0 else -noise)
        sd, at = {}, 0
        for name, w in module.state_dict().items():                     # policy.py:49-59
            n = w.numel()
            sd[name] = torch.from_numpy(np.reshape(params[at:at + n], w.shape))
            at += n
        module.load_state_dict(sd)
        rews = []
        with torch.no_grad():
            for t in range(env.T):                                       # gym_runner.py:50-54
                ob = torch.from_numpy(env.obs_stream[t]).float()
                x = torch.clamp((ob - _G['obmean']) / _G['obstd'], min=-5.0, max=5.0)
                a = module(x.float())
                rews.append(float(np.dot(a.numpy(), env.rew_vec[t])))
        out.append(sum(rews))
    return out


def _worker(args):
    import torch
    from oracle import es_oracle as orc
    torch.set_num_threads(1)
    rank, n_pairs, K_update, seed = args
    table, theta, env, dims, P = _G['table'], _G['theta'], _G['env'], _G['dims'], _G['P']
    module = _G.get('module')
    if module is None:
        layers = []
        for i, o in dims:
            layers += [torch.nn.Linear(i, o), torch.nn.Tanh()]
        module = _G['module'] = torch.nn.Sequential(*layers)
    rs = np.random.RandomState(seed)
    t0 = time.perf_counter()
    for _ in range(n_pairs):
        idx = orc.sample_idx(len(table), rs, P)
        rs.random(); rs.random()
        _reference_rollout_pair(table, theta, 0.02, module, dims, env, idx, P)
    t_roll = time.perf_counter() - t0
    # rank + reconstruct + Adam on K_update pairs (synthetic fitness of the right shape)
    r2 = np.random.RandomState(seed + 7)
    pos, neg = r2.randn(K_update, 1), r2.randn(K_update, 1)
    inds = r2.randint(0, len(table) - P, size=K_update).astype(np.float64)
    flat, opt = theta.copy(), orc.AdamOracle(P, 0.01)
    t1 = time.perf_counter()
    w, n_ranked = orc.centered_ranker(pos, neg)
    orc.approx_grad(flat, opt, w, inds, n_ranked, table, 500, 0.005)
    t_update = time.perf_counter() - t1
    return dict(rank=rank, pairs=n_pairs, t_roll=t_roll, t_update=t_update)


def host_cores() -> int:
    """Cores this process may actually use: the affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


class CpuReference:
    """Pool of one worker per host core sharing the table; ``sample()`` times one bounded
    sample of a K-pair generation and extrapolates (rollouts linear in pairs, update linear in K)."""

    def __init__(self, K: int, obs_dim: int, act_dim: int, hidden, T: int, table_len: int = 1 << 25,
                 cores: int = 0, seed: int = 1000):
        import torch
        from oracle import es_oracle as orc
        self.cores = cores or host_cores()
        self.K, self.seed = K, seed
        dims = orc.layer_dims(obs_dim, hidden, act_dim)
        P = orc.n_params(dims)
        table = np.random.default_rng(123).standard_normal(table_len, dtype=np.float32)   # content irrelevant to timing
        _G.update(table=table, theta=(np.random.RandomState(7).randn(P) * 0.1).astype(np.float32),
                  env=orc.SyntheticEnvSpec(obs_dim, act_dim, T), dims=dims, P=P,
                  obmean=torch.zeros(obs_dim, dtype=torch.float64), obstd=torch.ones(obs_dim, dtype=torch.float64))
        self.table_len = table_len
        self.pool = mp.get_context('fork').Pool(self.cores)
        self.versions = dict(numpy=np.__version__, torch=torch.__version__)

    def sample(self, pairs_per_worker: int = 1, K_update: int = 2000) -> dict:
        K_update = min(K_update, self.K)
        res = self.pool.map(_worker, [(r, pairs_per_worker, K_update, self.seed + r) for r in range(self.cores)])
        per_pair = max(r['t_roll'] / r['pairs'] for r in res)            # slowest rank sets the pace (MPI barrier)
        t_roll_full = per_pair * (self.K / self.cores)
        t_update_full = max(r['t_update'] for r in res) * (self.K / K_update)
        total = t_roll_full + t_update_full
        return dict(cores=self.cores, K=self.K, sec_per_pair_per_core=per_pair, t_rollouts_s=t_roll_full,
                    t_update_s=t_update_full, t_generation_s=total, pairs_per_sec=self.K / total,
                    sample=f'{pairs_per_worker} pair(s) x T per worker on {self.cores} worker processes, extrapolated '
                           f'linearly to K/{self.cores} pairs per rank; rank+scale_noise(batch 500)+Adam timed on '
                           f'{K_update} pairs in all workers at once, scaled by K/{K_update}; table {self.table_len} floats',
                    **self.versions)

    def close(self):
        self.pool.close()
        self.pool.join()


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--K', type=int, default=10000)
    ap.add_argument('--obs', type=int, default=376)
    ap.add_argument('--act', type=int, default=17)
    ap.add_argument('--T', type=int, default=1000)
    ap.add_argument('--pairs-per-worker', type=int, default=2)
    ap.add_argument('--cores', type=int, default=0)
    a = ap.parse_args()
    ref = CpuReference(a.K, a.obs, a.act, (64, 64), a.T, cores=a.cores)
    print(json.dumps(ref.sample(a.pairs_per_worker)))
    ref.close()
