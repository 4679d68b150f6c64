"""Generates tests/golden/ref_step.npz and tests/golden/policy-ref by running the REAL reference (/root/reference, build
container only) through its own ``es.step`` and ``Policy.save``:

  A. two generations of ``src.core.es.step`` (es.py:23-51) with an obj.py-style fit_fn that draws ``rs.random()`` in EVERY
     call -- including the noiseless ``fit_fn(policy.pheno(zeros), False)`` of es.py:48 -- recording theta, the noiseless
     result and the RandomState after each step (pins the RNG position a full reference step leaves behind);
  B. ``Policy.save`` of that policy (real ``src.core.policy.Policy`` pickle: module, flat_params, obstat, Adam m/v/t), then a
     third ``es.step`` of the RELOADED policy (``Policy.load``) -- what a resumed run must reproduce (obj.py, run_saved.py);
  C. one generation with action noise (``ac_std = 0.01``, the value of every shipped config; nn.py:47-48 draws
     ``rs.randn(act)`` from the same stream at every step of every rollout), simple_example.py's r_fn.

Same inert stand-ins for the absent third-party imports as make_ref_pipeline.py.  Nothing from /root/reference is copied:
it is imported and executed.

    python tests/golden/make_ref_step.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_ref_pipeline import REF, install_stand_ins  # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__


class QuietReporter:
    def print(self, s): pass
    def log(self, d): pass
    def log_gen(self, fits, noiseless_tr, policy, steps): pass


def main():
    comm = install_stand_ins()
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import torch
    from src.core import es
    from src.core.noisetable import NoiseTable
    from src.core.policy import Policy
    from src.gym import gym_runner
    from src.gym.training_result import RewardResult
    from src.nn.nn import FeedForward
    from src.nn.obstat import ObStat
    from src.nn.optimizers import Adam
    from src.utils.rankers import CenteredRanker
    from es_pytorch_b200.gym.synthetic_env import SyntheticEnv        # the synthetic env is this repo's, numpy only here

    obs_dim, act_dim, hidden, T, n_pairs = 17, 6, [64, 64], 40, 6
    env = SyntheticEnv(obs_dim, act_dim, T)
    torch.manual_seed(0)
    net = FeedForward(list(hidden), torch.nn.Tanh(), env, 0.0, 5)
    policy = Policy(net, 0.02, Adam(len(Policy.get_flat(net)), 0.01))
    P = len(policy)
    theta0 = (np.random.RandomState(6).randn(P) * 0.1).astype(np.float32)
    policy.flat_params = theta0.copy()
    table = np.random.RandomState(5).randn(200_003).astype(np.float32)
    nt = NoiseTable(P, table)
    rs = np.random.RandomState(7000)
    save_obs_chance = 0.3
    cfg = Cfg(general=Cfg(policies_per_gen=2 * n_pairs, batch_size=500), policy=Cfg(l2coeff=0.005))

    def r_fn(model, use_ac_noise=True):                         # obj.py:53-57's shape: the coin is drawn in every call
        save_obs = rs.random() < save_obs_chance
        rews, behv, obs, steps = gym_runner.run_model(model, env, T, rs if use_ac_noise else None)
        return RewardResult(rews, behv, obs if save_obs else np.array([np.zeros(env.observation_space.shape)]), steps)

    out = dict(theta0=theta0, table_seed=np.array(5), table_len=np.array(len(table)), cfg=np.array([obs_dim, act_dim, T, n_pairs]),
               hidden=np.array(hidden), save_obs_chance=np.array(save_obs_chance), seed=np.array(7000))
    ranker = CenteredRanker()
    for g in range(2):
        tr, gen_obstat = es.step(cfg, comm, policy, nt, env, r_fn, rs, ranker, QuietReporter())
        policy.update_obstat(gen_obstat)
        st = rs.get_state()
        out[f's{g}_theta'], out[f's{g}_noiseless'] = policy.flat_params.copy(), np.array(tr.result)
        out[f's{g}_rs_key'], out[f's{g}_rs_pos'] = st[1].copy(), np.array(st[2])
        out[f's{g}_fits'], out[f's{g}_inds'] = np.asarray(ranker.fits), np.asarray(ranker.noise_inds)
        out[f's{g}_ob_sum'], out[f's{g}_ob_count'] = gen_obstat.sum.copy(), np.array(gen_obstat.count)
    # ---- B: the reference's own checkpoint, and the step a resumed reference run takes from it ----
    policy.save(HERE, 'ref')                                    # -> tests/golden/policy-ref (policy.py:43-47)
    out['ckpt_m'], out['ckpt_v'], out['ckpt_t'] = policy.optim.m.copy(), policy.optim.v.copy(), np.array(policy.optim.t)
    out['ckpt_obstat_sum'], out['ckpt_obstat_sumsq'], out['ckpt_obstat_count'] = policy.obstat.sum.copy(), policy.obstat.sumsq.copy(), np.array(policy.obstat.count)
    resumed = Policy.load(os.path.join(HERE, 'policy-ref'))
    tr, gen_obstat = es.step(cfg, comm, resumed, nt, env, r_fn, rs, CenteredRanker(), QuietReporter())
    st = rs.get_state()
    out['s2_theta'], out['s2_noiseless'], out['s2_rs_key'], out['s2_rs_pos'] = resumed.flat_params.copy(), np.array(tr.result), st[1].copy(), np.array(st[2])
    # ---- C: action noise (ac_std = 0.01 as in configs/simple_conf.json:14, obj.json:18, nsra.json:17) ----
    torch.manual_seed(1)
    net_n = FeedForward(list(hidden), torch.nn.Tanh(), env, 0.01, 5)
    pol_n = Policy(net_n, 0.02, Adam(P, 0.01))
    pol_n.flat_params = theta0.copy()
    rs_n = np.random.RandomState(8000)

    def noisy_fn(model):                                        # simple_example.py:37-40: run_model gets the stream
        save_obs = rs_n.random() < save_obs_chance
        rews, behv, obs, steps = gym_runner.run_model(model, env, T, rs_n)
        return RewardResult(rews, behv, obs if save_obs else np.array([np.zeros(env.observation_space.shape)]), steps)

    gen_obstat = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(comm, n_pairs, pol_n, nt, gen_obstat, noisy_fn, rs_n)
    cr = CenteredRanker()
    ranked = cr.rank(pos, neg, inds)
    es.approx_grad(pol_n, cr, nt, pol_n.flat_params, 500, 0.005)
    st = rs_n.get_state()
    out.update(acn_std=np.array(0.01), acn_seed=np.array(8000), acn_pos=pos, acn_neg=neg, acn_inds=inds, acn_w=np.asarray(ranked),
               acn_theta=pol_n.flat_params.copy(), acn_rs_key=st[1].copy(), acn_rs_pos=np.array(st[2]),
               acn_rs_has_gauss=np.array(st[3]), acn_rs_gauss=np.array(st[4]),
               acn_ob_sum=gen_obstat.sum.copy(), acn_ob_count=np.array(gen_obstat.count))
    np.savez_compressed(os.path.join(HERE, 'ref_step.npz'), **out)
    print('ref_step.npz', len(out), 'arrays;', 'policy-ref', os.path.getsize(os.path.join(HERE, 'policy-ref')), 'bytes')


if __name__ == '__main__':
    main()
