"""Generates tests/golden/ref_pipeline.npz by running the REAL reference pipeline (/root/reference, build container only):

    src.core.es.test_params -> src.utils.rankers.CenteredRanker.rank -> src.core.es.approx_grad -> Policy.update_obstat

with the real NoiseTable / Policy / FeedForward / gym_runner.run_model / RewardResult / ObStat / Adam on the synthetic open-loop
env of SURVEY.md section 8d, for two generations.  This pins what no reference test pins: Policy.pheno, FeedForward.forward,
run_model, the RNG interleaving of test_params, approx_grad, the obs-statistics feedback.

The reference's third-party imports that are absent here (mpi4py, gym, munch, mlflow, mlagents_envs) are replaced by inert stand-ins defined
below (a 1-rank communicator whose collectives are identities, an empty gym namespace, a dict-with-attributes Munch, no-op
mlflow functions), and ``np.float`` (removed from numpy >= 1.24, used at src/core/es.py:89) is restored as ``float``.  None
of this touches the reference's arithmetic.  Nothing from /root/reference is copied: it is imported and executed.

    python tests/golden/make_ref_pipeline.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def install_stand_ins():
    np.float = float                                           # src/core/es.py:89
    mpi4py, MPI = types.ModuleType('mpi4py'), types.ModuleType('mpi4py.MPI')

    class Op:
        @staticmethod
        def Create(fn, commute=True):
            return fn

    class Comm:
        rank, size = 0, 1

        def Get_rank(self): return 0
        def Get_size(self): return 1
        def Alltoall(self, send, recv): recv[...] = send       # one rank: the exchange is the identity
        def alltoall(self, x): return list(x)
        def allreduce(self, x, op=None): return x
        def bcast(self, x, root=0): return x
        def gather(self, x, root=0): return [x]
        def scatter(self, x, root=0): return x[0]
        def Barrier(self): pass

    MPI.Op, MPI.Comm, MPI.COMM_WORLD, MPI.SUM, MPI.Win, MPI.FLOAT, MPI.COMM_TYPE_SHARED = Op, Comm, Comm(), 'sum', object, 4, 0
    mpi4py.MPI = MPI
    gym = types.ModuleType('gym')
    gym.Env = object
    gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('no gym here'))
    gym.utils = types.ModuleType('gym.utils')
    gym.utils.seeding = types.ModuleType('gym.utils.seeding')
    gym.utils.seeding._int_list_from_bigint = lambda x: [x]
    gym.utils.seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    gym.spaces = types.ModuleType('gym.spaces')
    munch = types.ModuleType('munch')

    class Munch(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    munch.Munch, munch.munchify, munch.unmunchify = Munch, (lambda d: d), (lambda d: d)
    mlflow = types.ModuleType('mlflow')
    for name in ('log_params', 'log_metrics', 'set_experiment', 'start_run'):
        setattr(mlflow, name, lambda *a, **k: None)
    # src/gym/gym_runner.py:8 imports the Unity wrapper, which imports mlagents_envs (absent): empty namespaces
    ml = {n: types.ModuleType(n) for n in ('mlagents_envs', 'mlagents_envs.base_env', 'mlagents_envs.environment',
                                           'mlagents_envs.side_channel',
                                           'mlagents_envs.side_channel.engine_configuration_channel')}
    ml['mlagents_envs.base_env'].ActionTuple = object
    ml['mlagents_envs.environment'].UnityEnvironment = object
    ml['mlagents_envs.side_channel.engine_configuration_channel'].EngineConfigurationChannel = object
    sys.modules.update(ml)
    for name, mod in (('mpi4py', mpi4py), ('mpi4py.MPI', MPI), ('gym', gym), ('gym.utils', gym.utils),
                      ('gym.utils.seeding', gym.utils.seeding), ('gym.spaces', gym.spaces), ('munch', munch), ('mlflow', mlflow)):
        sys.modules[name] = mod
    return MPI.COMM_WORLD


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
    rs = np.random.RandomState(1000)
    save_obs_chance = 0.3

    def r_fn(model):                                            # simple_example.py:37-40
        save_obs = rs.random() < save_obs_chance
        rews, behv, obs, steps = gym_runner.run_model(model, env, T, rs)
        return RewardResult(rews, behv, obs if save_obs else np.array([np.zeros(env.observation_space.shape)]), steps)

    out = dict(theta0=theta0, table_seed=np.array(5), table_len=np.array(len(table)), cfg=np.array([obs_dim, act_dim, T, n_pairs]),
               hidden=np.array(hidden), save_obs_chance=np.array(save_obs_chance), seed=np.array(1000))
    ranker = CenteredRanker()
    for g in range(2):
        gen_obstat = ObStat(env.observation_space.shape, 0)
        pos, neg, inds, steps = es.test_params(comm, n_pairs, policy, nt, gen_obstat, r_fn, rs)
        out[f'g{g}_obmean'], out[f'g{g}_obstd'] = np.array(net._obmean, dtype=np.float64), np.array(net._obstd, dtype=np.float64)
        policy.update_obstat(gen_obstat)
        ranked = ranker.rank(pos, neg, inds)
        es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
        out[f'g{g}_pos'], out[f'g{g}_neg'], out[f'g{g}_inds'], out[f'g{g}_steps'] = pos, neg, inds, np.array(steps)
        out[f'g{g}_w'], out[f'g{g}_n_ranked'] = np.asarray(ranked), np.array(ranker.n_fits_ranked)
        out[f'g{g}_ob_sum'], out[f'g{g}_ob_sumsq'], out[f'g{g}_ob_count'] = gen_obstat.sum, gen_obstat.sumsq, np.array(gen_obstat.count)
        out[f'g{g}_theta'] = policy.flat_params.copy()
    out['rs_key'], out['rs_pos'] = rs.get_state()[1], np.array(rs.get_state()[2])
    # one noiseless evaluation (es.py:48) of the final policy
    tr = r_fn(policy.pheno(np.zeros(len(policy))))
    out['noiseless_result'], out['noiseless_behv'] = np.array(tr.result), np.array(tr.behaviour)
    # ---- NSRA-style generation: real NSRResult (reward + novelty of the final (x, y)) and MultiObjectiveRanker ----
    from src.gym.training_result import NSRResult
    from src.utils.rankers import EliteRanker, MultiObjectiveRanker
    archive = np.random.RandomState(17).randn(16, 2)
    policy.flat_params = theta0.copy()
    policy.optim = Adam(P, 0.01)
    net.set_ob_mean_std(np.zeros(obs_dim), np.ones(obs_dim))
    rs2 = np.random.RandomState(2000)

    def nsr_fn(model):                                          # nsra.py's fit_fn shape: one coin, then the rollout
        save_obs = rs2.random() < save_obs_chance
        rews, behv, obs, steps = gym_runner.run_model(model, env, T, rs2)
        return NSRResult(rews, behv, obs if save_obs else np.array([np.zeros(env.observation_space.shape)]), steps, archive, 10)

    gen_obstat = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(comm, n_pairs, policy, nt, gen_obstat, nsr_fn, rs2)
    moo = MultiObjectiveRanker(CenteredRanker(), 0.5)
    ranked = moo.rank(pos, neg, inds)
    es.approx_grad(policy, moo, nt, policy.flat_params, 500, 0.005)
    out.update(nsra_archive=archive, nsra_seed=np.array(2000), nsra_pos=pos, nsra_neg=neg, nsra_inds=inds,
               nsra_w=np.asarray(ranked), nsra_theta=policy.flat_params.copy())
    # ---- obj.py's EliteRanker(CenteredRanker(), elite) through the real approx_grad ----
    policy.flat_params = theta0.copy()
    policy.optim = Adam(P, 0.01)
    rs3 = np.random.RandomState(3000)

    def r3(model):
        rs3.random()
        rews, behv, obs, steps = gym_runner.run_model(model, env, T, rs3)
        return RewardResult(rews, behv, np.array([np.zeros(env.observation_space.shape)]), steps)

    gen_obstat = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(comm, n_pairs, policy, nt, gen_obstat, r3, rs3)
    elite = EliteRanker(CenteredRanker(), 0.25)
    vals = np.asarray(elite.rank(pos, neg, inds))
    es.approx_grad(policy, elite, nt, policy.flat_params, 500, 0.005)
    order = np.lexsort((elite.noise_inds, vals))
    out.update(elite_seed=np.array(3000), elite_pct=np.array(0.25), elite_pos=pos, elite_neg=neg, elite_inds=inds,
               elite_vals=vals[order], elite_sel=np.asarray(elite.noise_inds)[order], elite_n=np.array(elite.n_fits_ranked),
               elite_theta=policy.flat_params.copy())
    # ---- the other optimizers through the real approx_grad (momentum SGD over two updates, SimpleES) ----
    from src.nn.optimizers import SGD, SimpleES
    for tag, make in (('sgd', lambda: SGD(P, 0.01)), ('simple', lambda: SimpleES(P, 0.01))):
        policy.flat_params = theta0.copy()
        policy.optim = make()
        rs4 = np.random.RandomState(5000)

        def r4(model):
            rs4.random()
            rews, behv, obs, steps = gym_runner.run_model(model, env, T, rs4)
            return RewardResult(rews, behv, np.array([np.zeros(env.observation_space.shape)]), steps)

        for g in range(2):
            pos, neg, inds, steps = es.test_params(comm, n_pairs, policy, nt, ObStat(env.observation_space.shape, 0), r4, rs4)
            cr = CenteredRanker()
            cr.rank(pos, neg, inds)
            es.approx_grad(policy, cr, nt, policy.flat_params, 500, 0.005)
            out[f'{tag}_g{g}_inds'], out[f'{tag}_g{g}_theta'] = inds, policy.flat_params.copy()
    # ---- the bench's policy shape (Humanoid-shaped 376-64-64-17), short episode, one generation ----
    env_h = SyntheticEnv(376, 17, 16)
    net_h = FeedForward([64, 64], torch.nn.Tanh(), env_h, 0.0, 5)
    Ph = len(Policy.get_flat(net_h))
    pol_h = Policy(net_h, 0.02, Adam(Ph, 0.01))
    theta_h = (np.random.RandomState(8).randn(Ph) * 0.1).astype(np.float32)
    pol_h.flat_params = theta_h.copy()
    nt_h = NoiseTable(Ph, table)
    rs5 = np.random.RandomState(6000)

    def r5(model):
        rs5.random()
        rews, behv, obs, steps = gym_runner.run_model(model, env_h, 16, rs5)
        return RewardResult(rews, behv, np.array([np.zeros(env_h.observation_space.shape)]), steps)

    pos, neg, inds, steps = es.test_params(comm, 3, pol_h, nt_h, ObStat(env_h.observation_space.shape, 0), r5, rs5)
    cr = CenteredRanker()
    ranked = cr.rank(pos, neg, inds)
    es.approx_grad(pol_h, cr, nt_h, pol_h.flat_params, 500, 0.005)
    out.update(hum_theta0=theta_h, hum_pos=pos, hum_neg=neg, hum_inds=inds, hum_w=np.asarray(ranked), hum_theta=pol_h.flat_params.copy())
    # ---- two MPI ranks: the real test_params on two threads, each with its own Policy / RandomState, joined by a
    #      communicator whose Alltoall / allreduce do what MPI's would for size 2 (pins the rank-major layout of
    #      es._share_results, the per-rank RNG streams and ObStat.mpi_inc) ----
    import threading

    class TwoRankWorld:
        def __init__(self):
            self.barrier = threading.Barrier(2)
            self.slots = [None, None]

        def exchange(self, rank, value):
            self.slots[rank] = value
            self.barrier.wait()
            both = list(self.slots)
            self.barrier.wait()
            return both

    class RankComm:
        size = 2

        def __init__(self, world, rank):
            self.world, self.rank = world, rank

        def Get_rank(self): return self.rank
        def Get_size(self): return 2

        def Alltoall(self, send, recv):                        # block j of recv <- block `rank` of rank j's send
            both = self.world.exchange(self.rank, np.array(send, copy=True))
            n = send.shape[0] // 2
            for j in range(2):
                recv[j * n:(j + 1) * n] = both[j][self.rank * n:(self.rank + 1) * n]

        def allreduce(self, x, op=None):
            import copy
            both = self.world.exchange(self.rank, copy.deepcopy(x))
            if callable(op):                                   # MPI.Op.Create(sum_obstat): fold in rank order
                acc = copy.deepcopy(both[0])
                return op(acc, both[1], None)
            return both[0] + both[1]

    world = TwoRankWorld()
    results = [None, None]

    def rank_main(rank):
        torch.manual_seed(rank)
        env_r = SyntheticEnv(obs_dim, act_dim, T)
        net_r = FeedForward(list(hidden), torch.nn.Tanh(), env_r, 0.0, 5)
        pol_r = Policy(net_r, 0.02, Adam(P, 0.01))
        pol_r.flat_params = theta0.copy()
        rs_r = np.random.RandomState(4000 + rank)

        def fn(model):
            save_obs = rs_r.random() < save_obs_chance
            rews, behv, obs, steps = gym_runner.run_model(model, env_r, T, rs_r)
            return RewardResult(rews, behv, obs if save_obs else np.array([np.zeros(env_r.observation_space.shape)]), steps)

        st = ObStat(env_r.observation_space.shape, 0)
        results[rank] = es.test_params(RankComm(world, rank), 4, pol_r, nt, st, fn, rs_r) + (st,)

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    (p0, n0, i0, s0, st0), (p1, n1, i1, s1, st1) = results
    assert np.array_equal(p0, p1) and np.array_equal(i0, i1) and s0 == s1 and np.array_equal(st0.sum, st1.sum)   # every rank sees all
    out.update(two_seeds=np.array([4000, 4001]), two_n=np.array(4), two_pos=p0, two_neg=n0, two_inds=i0, two_steps=np.array(s0),
               two_ob_sum=st0.sum, two_ob_sumsq=st0.sumsq, two_ob_count=np.array(st0.count))
    np.savez_compressed(os.path.join(HERE, 'ref_pipeline.npz'), **out)
    print('ref_pipeline.npz', len(out), 'arrays; theta dtype', out['g1_theta'].dtype, 'pos dtype', out['g0_pos'].dtype)


if __name__ == '__main__':
    main()
