"""A training script written against the REFERENCE's import surface only -- ``src.*``, ``mpi4py.MPI.COMM_WORLD``,
``gym.make``, ``utils.load_config`` with the schema of configs/simple_conf.json / obj.json -- in the two shapes the
reference's own scripts have:

  explicit   the generation spelled out (simple_example.py:45-58): an opaque per-policy fit_fn that draws its save_obs coin
             and calls gym_runner.run_model with the rank's RandomState, es.test_params, policy.update_obstat, ranker.rank,
             es.approx_grad;
  step       es.step per generation with a fit_fn(model, use_ac_noise) and the between-generation decays of obj.py:77-83.

Run with PYTHONPATH=<repo>:<repo>/es_pytorch_b200/compat (the shims that stand in for mpi4py / gym / munch / mlflow):

    python tests/drivers/es_script.py <explicit|step> <config.json> <out_prefix>
    torchrun --nproc-per-node 2 ... tests/drivers/es_script.py ...

Every rank writes <out_prefix>.rank<r>.npz with what a parity check needs (initial theta, per-generation theta / fits /
indices / stream state); tests/test_gpu_scripts.py replays the same run with the CPU oracle."""
import sys

import gym
import numpy as np
import torch
from mpi4py import MPI

import src.core.es as es
from src.core.noisetable import NoiseTable
from src.core.policy import Policy
from src.gym import gym_runner
from src.gym.training_result import RewardResult
from src.nn.nn import FeedForward
from src.nn.obstat import ObStat
from src.nn.optimizers import Adam
from src.utils import utils
from src.utils.rankers import CenteredRanker
from src.utils.reporters import StdoutReporter


def main(mode, cfg_file, out_prefix):
    comm = MPI.COMM_WORLD
    cfg = utils.load_config(cfg_file)
    env = gym.make(cfg.env.name)
    rs, my_seed, global_seed = utils.seed(comm, cfg.general.seed, env)
    net = FeedForward(cfg.policy.layer_sizes, torch.nn.Tanh(), env, cfg.policy.ac_std, cfg.policy.ob_clip)
    policy = Policy(net, cfg.noise.std, Adam(len(Policy.get_flat(net)), cfg.policy.lr))
    nt = NoiseTable.create_shared(comm, cfg.noise.tbl_size, len(policy), None, cfg.general.seed)
    ranker = CenteredRanker()
    reporter = StdoutReporter(comm)
    max_steps = int(cfg.env.get('max_steps', 10000))
    no_obs = np.array([np.zeros(env.observation_space.shape)])
    log = dict(theta0=policy.flat_params.copy(), my_seed=np.array(my_seed), global_seed=np.array(global_seed))

    def fit_fn(model, use_ac_noise=True):
        save_obs = rs.random() < cfg.policy.save_obs_chance
        rews, behv, obs, steps = gym_runner.run_model(model, env, max_steps, rs if use_ac_noise else None)
        return RewardResult(rews, behv, obs if save_obs else no_obs, steps)

    pairs_per_rank = int((cfg.general.policies_per_gen / comm.size) / 2)
    for gen in range(cfg.general.gens):
        if mode == 'explicit':
            gen_obstat = ObStat(env.observation_space.shape, 0)
            pos, neg, inds, steps = es.test_params(comm, pairs_per_rank, policy, nt, gen_obstat, fit_fn, rs)
            policy.update_obstat(gen_obstat)
            ranker.rank(pos, neg, inds)
            es.approx_grad(policy, ranker, nt, policy.flat_params, cfg.general.batch_size, cfg.policy.l2coeff)
        else:
            tr, gen_obstat = es.step(cfg, comm, policy, nt, env, fit_fn, rs, ranker, reporter)
            policy.update_obstat(gen_obstat)
            log[f'g{gen}_noiseless'] = np.asarray(tr.result, dtype=np.float64)
            cfg.policy.ac_std = net._action_std = net._action_std * cfg.policy.ac_std_decay
            cfg.noise.std = policy.std = max(cfg.noise.std * cfg.noise.std_decay, cfg.noise.std_limit)
            cfg.policy.lr = policy.optim.lr = max(cfg.policy.lr * cfg.policy.lr_decay, cfg.policy.lr_limit)
        st = rs.get_state()
        log[f'g{gen}_theta'] = policy.flat_params.copy()
        log[f'g{gen}_fits'] = np.asarray(ranker.fits, dtype=np.float64)
        log[f'g{gen}_inds'] = np.asarray(ranker.noise_inds, dtype=np.float64)
        log[f'g{gen}_rs_key'], log[f'g{gen}_rs_pos'] = st[1].copy(), np.array(st[2])
        log[f'g{gen}_rs_has_gauss'], log[f'g{gen}_rs_gauss'] = np.array(st[3]), np.array(st[4])
        log[f'g{gen}_ob_count'] = np.array(float(policy.obstat.count))
        if comm.rank == 0:
            print(f'gen {gen}: mean fitness {float(np.mean(ranker.fits)):.4f}', flush=True)
    np.savez(f'{out_prefix}.rank{comm.rank}.npz', **log)
    if comm.rank == 0:
        print('SCRIPT_DONE', flush=True)


if __name__ == '__main__':
    main(*sys.argv[1:4])
