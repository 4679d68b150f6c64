"""GPU parity of whole generations: the device pipeline (DeviceGeneration) and the
reference-facing API (es.test_params / Ranker.rank / es.approx_grad / es.step) against the
oracle's restatement of src/core/es.py:38-101 on the same seeds, plus size-independent
properties at BASELINE.json's full sizes."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import es_oracle as orc

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))


def _small(eng):
    from make_golden import small_problem
    dims, P, table, theta, env = small_problem()
    return dims, P, table, theta, env


def _mk_generation(eng, table, theta, env, streams, optim, **kw):
    from es_pytorch_b200.generation import DeviceGeneration
    sizes = kw.pop('sizes')
    return DeviceGeneration(eng.to_device(table), eng.to_device(theta.copy()), sizes, eng.to_device(env.obs_stream),
                            eng.to_device(env.rew_vec), streams, 0.02, 0.005, optim, ob_clip=5.0,
                            pos_scale=env.pos_scale, engine=eng, **kw)


def test_device_generation_matches_golden(eng, oracle_vectors):
    """Two generations, 2 virtual ranks x 6 pairs, one save_obs coin per evaluation, Adam."""
    from es_pytorch_b200.nn.optimizers import Adam
    v = oracle_vectors
    dims, P, table, theta, env = _small(eng)
    streams = [np.random.RandomState(1000), np.random.RandomState(1001)]
    gen = _mk_generation(eng, table, theta, env, streams, Adam(P, 0.01), sizes=[17, 64, 64, 6], coins_per_eval=1,
                         save_obs_chance=0.0)
    gen.set_obstat(v['obmean'], v['obstd'])
    for g in range(2):
        fpos, fneg = gen.evaluate(6)
        assert np.array_equal(gen.idx.cpu().numpy(), v[f'gen{g}_inds'].astype(np.int64))     # indices bit-exact
        scale = 40.0
        assert np.abs(fpos.cpu().numpy() - v[f'gen{g}_pos']).max() <= 1e-5 * scale
        assert np.abs(fneg.cpu().numpy() - v[f'gen{g}_neg']).max() <= 1e-5 * scale
        gen.update(fpos, fneg)
        assert np.array_equal(gen.weights.cpu().numpy(), v[f'gen{g}_w'])                       # rank weights bit-exact
        # theta after the Adam step: gradient within 1e-5 rel -> theta within a few ulp of the step size
        assert np.abs(gen.theta.cpu().numpy() - v[f'gen{g}_theta']).max() <= 2e-6
    # the device streams are where numpy's would be after the same draws
    ref = [np.random.RandomState(1000), np.random.RandomState(1001)]
    for r in ref:
        for _ in range(12):
            r.randint(0, len(table) - P); r.random(); r.random()
    for a, b in zip(gen.rank_states(), ref):
        assert np.array_equal(a.get_state()[1], b.get_state()[1]) and a.get_state()[2] == b.get_state()[2]


@pytest.mark.parametrize('shaping,elite', [('centered', 0.25), ('double_positive', None), ('semi_centered', None),
                                           ('max_normalized', None), ('max_normalized', 0.5)])
def test_device_generation_other_rankers(eng, oracle_vectors, shaping, elite):
    """DeviceGeneration(ranker=...) with the rankers of rankers.py:61-103 (obj.py:50 uses the elite one) against the
    oracle's generation on the same seeds."""
    from es_pytorch_b200.nn.optimizers import Adam
    from es_pytorch_b200.utils import rankers as R
    v = oracle_vectors
    dims, P, table, theta, env = _small(eng)
    cls = {'centered': R.CenteredRanker, 'double_positive': R.DoublePositiveCenteredRanker,
           'semi_centered': R.SemiCenteredRanker, 'max_normalized': R.MaxNormalizedRanker}[shaping]
    ranker = cls() if elite is None else R.EliteRanker(cls(), elite)
    streams = [np.random.RandomState(1000), np.random.RandomState(1001)]
    gen = _mk_generation(eng, table, theta, env, streams, Adam(P, 0.01), sizes=[17, 64, 64, 6], ranker=ranker)
    gen.set_obstat(v['obmean'], v['obstd'])
    flat, opt = theta.copy(), orc.AdamOracle(P, 0.01)
    ostates = [np.random.RandomState(1000), np.random.RandomState(1001)]
    for g in range(2):
        res = orc.generation(table, flat, opt, 0.02, dims, env, [1000, 1001], 6, v['obmean'], v['obstd'], 5.0, env.T,
                             500, 0.005, rank_states=ostates, shaping=shaping, elite_percent=elite)
        gen.run(6)
        assert ranker.n_fits_ranked == res['n_ranked']
        assert np.array_equal(gen.idx.cpu().numpy(), res['inds'].astype(np.int64))
        assert np.abs(gen.theta.cpu().numpy() - flat).max() <= 2e-6


def test_device_generation_nsra_matches_golden(eng, oracle_vectors):
    from es_pytorch_b200.nn.optimizers import Adam
    v = oracle_vectors
    dims, P, table, theta, env = _small(eng)
    archive = np.random.RandomState(17).randn(16, 2)
    gen = _mk_generation(eng, table, theta, env, [np.random.RandomState(1000), np.random.RandomState(1001)],
                         Adam(P, 0.01), sizes=[17, 64, 64, 6], coins_per_eval=1,
                         archive=eng.to_device(archive), nov_k=10, moo_w=0.5)
    gen.set_obstat(v['obmean'], v['obstd'])
    fpos, fneg = gen.evaluate(6)
    fp, fn = fpos.cpu().numpy(), fneg.cpu().numpy()
    assert np.abs(fp[:, 0] - v['nsra_pos'][:, 0]).max() <= 4e-4 and np.abs(fn[:, 0] - v['nsra_neg'][:, 0]).max() <= 4e-4
    assert np.allclose(fp[:, 1], v['nsra_pos'][:, 1], rtol=1e-5) and np.allclose(fn[:, 1], v['nsra_neg'][:, 1], rtol=1e-5)
    gen.update(fpos, fneg)
    assert np.array_equal(gen.weights.cpu().numpy(), v['nsra_w'])
    assert np.abs(gen.theta.cpu().numpy() - v['nsra_theta']).max() <= 2e-6


def test_generation_vs_live_oracle_with_obstat(eng):
    """Humanoid-shaped, 4 virtual ranks x 8 pairs, save_obs_chance 0.3 so the obs statistics path runs."""
    from es_pytorch_b200.nn.optimizers import SGD
    rs = np.random.RandomState(77)
    obs_dim, act_dim, hidden, T = 376, 17, (64, 64), 25
    dims = orc.layer_dims(obs_dim, hidden, act_dim)
    P = orc.n_params(dims)
    table = rs.randn(P + 400_000).astype(np.float32)
    theta = (rs.randn(P) * 0.05).astype(np.float32)
    env = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    seeds = [1000, 1001, 1002, 1003]
    gen = _mk_generation(eng, table, theta, env, [np.random.RandomState(s) for s in seeds], SGD(P, 0.01),
                         sizes=[obs_dim, 64, 64, act_dim], coins_per_eval=1, save_obs_chance=0.3)
    gen.run(8)
    flat, opt = theta.copy(), orc.SGDOracle(P, 0.01)
    res = orc.generation(table, flat, opt, 0.02, dims, env, seeds, 8, np.zeros(obs_dim), np.ones(obs_dim), 5.0, T, 500,
                         0.005, coins_per_eval=1, batched=True)
    # save_obs_chance only matters for obstat; replay it for the oracle
    pos, neg, inds, steps, obstat = orc.es_test_params(table, theta, 0.02, dims, env, seeds, 8, np.zeros(obs_dim),
                                                       np.ones(obs_dim), 5.0, T, coins_per_eval=1, save_obs_chance=0.3)
    assert np.array_equal(gen.idx.cpu().numpy(), res['inds'].astype(np.int64))
    assert np.array_equal(gen.weights.cpu().numpy(), res['weights'])
    assert np.abs(gen.theta.cpu().numpy() - flat).max() <= 2e-6
    assert obstat.count > 0
    assert np.array_equal(gen.gen_sum.cpu().numpy(), obstat.sum) and np.array_equal(gen.gen_sumsq.cpu().numpy(), obstat.sumsq)
    assert gen.gen_count.cpu().numpy()[0] == obstat.count


# ---- the reference-facing API -----------------------------------------------------------------------------
class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _api_objects(eng, table, theta, env_spec, hidden):
    from es_pytorch_b200.core.noisetable import NoiseTable
    from es_pytorch_b200.core.policy import Policy
    from es_pytorch_b200.gym.synthetic_env import SyntheticEnv
    from es_pytorch_b200.nn.nn import FeedForward
    from es_pytorch_b200.nn.optimizers import Adam
    env = SyntheticEnv(env_spec.obs_dim, env_spec.act_dim, env_spec.T)
    assert np.array_equal(env.obs_stream, env_spec.obs_stream) and np.array_equal(env.rew_vec, env_spec.rew_vec)
    net = FeedForward(list(hidden), torch.nn.Tanh(), env, 0.0, 5)
    policy = Policy(net, 0.02, Adam(len(theta), 0.01))
    policy.flat_params[...] = theta                        # parity harness passes theta in explicitly
    policy.set_nn_params(policy.flat_params)
    return env, net, policy, NoiseTable(len(theta), table)


def test_api_step_batched_matches_oracle(eng, oracle_vectors):
    """es.step with a BatchedRollout fit_fn == two oracle generations (same golden vectors as above)."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.utils.rankers import CenteredRanker
    from es_pytorch_b200.utils.reporters import ReporterSet
    v = oracle_vectors
    dims, P, table, theta, spec = _small(eng)
    env, net, policy, nt = _api_objects(eng, table, theta, spec, (64, 64))
    net.set_ob_mean_std(v['obmean'], v['obstd'])
    streams = [np.random.RandomState(1000), np.random.RandomState(1001)]
    fit_fn = BatchedRollout(env, spec.T, coins_per_eval=1, save_obs_chance=0.0, rank_streams=streams)
    cfg = _Cfg(general=_Cfg(policies_per_gen=12, batch_size=500), policy=_Cfg(l2coeff=0.005))
    comm = dist.world()
    ranker = CenteredRanker()
    for g in range(2):
        # 2 virtual ranks live in this one process: n = 6 pairs per stream
        gen_obstat_shape = env.observation_space.shape
        from es_pytorch_b200.nn.obstat import ObStat
        gen_obstat = ObStat(gen_obstat_shape, 0)
        pos, neg, inds, steps = es.test_params(comm, 6, policy, nt, gen_obstat, fit_fn, streams[0])
        assert pos.shape == (12, 1) and pos.dtype == np.float64 and inds.dtype == np.float64
        assert np.array_equal(inds, v[f'gen{g}_inds']) and steps == int(v[f'gen{g}_steps'])
        ranked = ranker.rank(pos, neg, inds)
        assert ranked.dtype == np.float32 and np.array_equal(ranked, v[f'gen{g}_w']) and ranker.n_fits_ranked == 24
        es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
        assert np.abs(policy.flat_params - v[f'gen{g}_theta']).max() <= 2e-6
    # callers' RandomState objects were advanced exactly like the reference would have
    ref = np.random.RandomState(1000)
    for _ in range(12):
        ref.randint(0, len(table) - P); ref.random(); ref.random()
    assert np.array_equal(streams[0].get_state()[1], ref.get_state()[1])
    # es.step end to end (one more generation): returns the noiseless TrainingResult and the generation ObStat
    tr, ob = es.step(cfg, comm, policy, nt, env, BatchedRollout(env, spec.T, coins_per_eval=1, rank_streams=None),
                     streams[0], CenteredRanker(), ReporterSet())
    assert len(tr.result) == 1 and np.isfinite(tr.result[0]) and ob.count == 0


def test_api_matches_real_reference_pipeline(eng):
    """The reference-facing API on the device against vectors produced by the REAL reference code
    (tests/golden/make_ref_pipeline.py ran src.core.es.test_params -> CenteredRanker.rank -> es.approx_grad ->
    Policy.update_obstat in the build container): noise indices, steps, obs statistics and rank weights bit-exact, fitness
    within the float32 rollout's tolerance, theta within 2e-6, the callers' RandomState where the reference left it."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.utils.rankers import CenteredRanker
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    env, net, policy, nt = _api_objects(eng, table, v['theta0'], spec, tuple(int(h) for h in v['hidden']))
    rs = np.random.RandomState(int(v['seed']))
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']))
    ranker, comm = CenteredRanker(), dist.world()
    for g in range(2):
        assert np.array_equal(net._obmean, v[f'g{g}_obmean']) and np.array_equal(net._obstd, v[f'g{g}_obstd'])
        gen_obstat = ObStat(env.observation_space.shape, 0)
        pos, neg, inds, steps = es.test_params(comm, n_pairs, policy, nt, gen_obstat, fit_fn, rs)
        assert np.array_equal(inds, v[f'g{g}_inds']) and steps == int(v[f'g{g}_steps'])
        scale = max(1.0, float(np.abs(v[f'g{g}_pos']).max())) * T ** 0.5
        assert np.abs(pos - v[f'g{g}_pos']).max() <= 1e-5 * scale and np.abs(neg - v[f'g{g}_neg']).max() <= 1e-5 * scale
        assert np.array_equal(gen_obstat.sum, v[f'g{g}_ob_sum']) and np.array_equal(gen_obstat.sumsq, v[f'g{g}_ob_sumsq'])
        assert gen_obstat.count == float(v[f'g{g}_ob_count'])
        policy.update_obstat(gen_obstat)
        ranked = ranker.rank(pos, neg, inds)
        assert np.array_equal(ranked, v[f'g{g}_w']) and ranker.n_fits_ranked == int(v[f'g{g}_n_ranked'])
        es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
        assert np.abs(policy.flat_params - v[f'g{g}_theta']).max() <= 2e-6
    assert np.array_equal(rs.get_state()[1], v['rs_key']) and rs.get_state()[2] == int(v['rs_pos'])
    # the noiseless evaluation of the final policy (es.py:48) as the reference's fit_fn returned it
    tr = fit_fn(policy.pheno(np.zeros(len(policy))), False)
    assert abs(tr.result[0] - float(v['noiseless_result'][0])) <= 1e-5 * max(1.0, abs(float(v['noiseless_result'][0]))) * T ** 0.5
    assert np.allclose(tr.behaviour, v['noiseless_behv'], rtol=1e-4, atol=1e-5)


def test_api_matches_real_reference_nsra_and_elite(eng):
    """Same vectors file: the real NSRResult + MultiObjectiveRanker generation and the real EliteRanker update."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.utils.rankers import CenteredRanker, EliteRanker, MultiObjectiveRanker
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    comm = dist.world()
    # NSRA
    env, net, policy, nt = _api_objects(eng, table, v['theta0'], spec, tuple(int(h) for h in v['hidden']))
    rs = np.random.RandomState(int(v['nsra_seed']))
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']), archive=v['nsra_archive'], nov_k=10)
    pos, neg, inds, _ = es.test_params(comm, n_pairs, policy, nt, ObStat(env.observation_space.shape, 0), fit_fn, rs)
    assert np.array_equal(inds, v['nsra_inds']) and pos.shape == (n_pairs, 2)
    scale = max(1.0, float(np.abs(v['nsra_pos']).max())) * T ** 0.5
    assert np.abs(pos - v['nsra_pos']).max() <= 1e-5 * scale and np.abs(neg - v['nsra_neg']).max() <= 1e-5 * scale
    moo = MultiObjectiveRanker(CenteredRanker(), 0.5)
    assert np.array_equal(moo.rank(pos, neg, inds), v['nsra_w'])
    es.approx_grad(policy, moo, nt, policy.flat_params, 500, 0.005)
    assert np.abs(policy.flat_params - v['nsra_theta']).max() <= 2e-6
    # Elite (obj.py:50)
    env, net, policy, nt = _api_objects(eng, table, v['theta0'], spec, tuple(int(h) for h in v['hidden']))
    rs = np.random.RandomState(int(v['elite_seed']))
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=0.0)
    pos, neg, inds, _ = es.test_params(comm, n_pairs, policy, nt, ObStat(env.observation_space.shape, 0), fit_fn, rs)
    assert np.array_equal(inds, v['elite_inds'])
    elite = EliteRanker(CenteredRanker(), float(v['elite_pct']))
    vals = np.asarray(elite.rank(pos, neg, inds))
    order = np.lexsort((elite.noise_inds, vals))
    assert elite.n_fits_ranked == int(v['elite_n']) and np.array_equal(vals[order], v['elite_vals'])
    assert np.array_equal(np.asarray(elite.noise_inds)[order], v['elite_sel'])
    es.approx_grad(policy, elite, nt, policy.flat_params, 500, 0.005)
    assert np.abs(policy.flat_params - v['elite_theta']).max() <= 2e-6


@pytest.mark.parametrize('tag', ['sgd', 'simple'])
def test_api_matches_real_reference_other_optimizers(eng, tag):
    """Momentum SGD (two consecutive updates) and SimpleES through es.approx_grad against the real reference's vectors."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.core.policy import Policy
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.nn.optimizers import SGD, SimpleES
    from es_pytorch_b200.utils.rankers import CenteredRanker
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    env, net, policy, nt = _api_objects(eng, table, v['theta0'], spec, tuple(int(h) for h in v['hidden']))
    P = len(v['theta0'])
    policy = Policy(net, 0.02, SGD(P, 0.01) if tag == 'sgd' else SimpleES(P, 0.01))
    policy.flat_params[...] = v['theta0']
    rs = np.random.RandomState(5000)
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=0.0)
    for g in range(2):
        pos, neg, inds, _ = es.test_params(dist.world(), n_pairs, policy, nt, ObStat(env.observation_space.shape, 0), fit_fn, rs)
        assert np.array_equal(inds, v[f'{tag}_g{g}_inds'])
        ranker = CenteredRanker()
        ranker.rank(pos, neg, inds)
        es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
        assert np.abs(policy.flat_params - v[f'{tag}_g{g}_theta']).max() <= 2e-6


@pytest.mark.parametrize('mode', ['f32', 'tc3', 'tc'])
def test_api_matches_real_reference_humanoid_shape(eng, mode):
    """The bench's policy shape (376-64-64-17) against the real reference's vectors: the float32 rollout AND the split
    tensor-core rollout (ES_ROLLOUT_TC3: float16 shadows + TMA, obs % 8 == 0) to the float32 tolerance -- same rank weights,
    theta within 2e-6 --; the single-product float16 rollout to its own (looser) tolerance; indices exact in every mode."""
    from es_pytorch_b200 import _lib, dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.utils.rankers import CenteredRanker
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    spec = orc.SyntheticEnvSpec(376, 17, 16)
    env, net, policy, nt = _api_objects(eng, table, v['hum_theta0'], spec, (64, 64))
    rs = np.random.RandomState(6000)
    fit_fn = BatchedRollout(env, 16, coins_per_eval=1, save_obs_chance=0.0,
                            rollout_mode={'f32': _lib.ES_ROLLOUT_F32, 'tc3': _lib.ES_ROLLOUT_TC3, 'tc': _lib.ES_ROLLOUT_TC}[mode])
    pos, neg, inds, _ = es.test_params(dist.world(), 3, policy, nt, ObStat(env.observation_space.shape, 0), fit_fn, rs)
    assert np.array_equal(inds, v['hum_inds'])
    ref = np.concatenate((v['hum_pos'], v['hum_neg'])).ravel()
    got = np.concatenate((pos, neg)).ravel()
    if mode in ('f32', 'tc3'):
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()) * 4.0
        ranker = CenteredRanker()
        assert np.array_equal(ranker.rank(pos, neg, inds), v['hum_w'])
        es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
        assert np.abs(policy.flat_params - v['hum_theta']).max() <= 2e-6
    else:
        spread = max(ref.std(), 1e-3 * 4.0)
        assert np.abs(got - ref).max() <= 0.02 * spread + 1e-3 * 4.0 * 0.05


def test_api_virtual_ranks_match_real_reference_two_ranks(eng):
    """One process carrying two RandomState streams ('virtual ranks') == the real reference on two MPI ranks (thread-emulated in
    make_ref_pipeline.py): rank-major indices and fitness rows, summed steps, merged obs statistics."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    obs_dim, act_dim, T, _ = [int(x) for x in v['cfg']]
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    env, net, policy, nt = _api_objects(eng, table, v['theta0'], spec, tuple(int(h) for h in v['hidden']))
    streams = [np.random.RandomState(int(s)) for s in v['two_seeds']]
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']), rank_streams=streams)
    st = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(dist.world(), int(v['two_n']), policy, nt, st, fit_fn, streams[0])
    assert np.array_equal(inds, v['two_inds']) and steps == int(v['two_steps'])
    scale = max(1.0, float(np.abs(v['two_pos']).max())) * T ** 0.5
    assert np.abs(pos - v['two_pos']).max() <= 1e-5 * scale and np.abs(neg - v['two_neg']).max() <= 1e-5 * scale
    assert np.array_equal(st.sum, v['two_ob_sum']) and np.array_equal(st.sumsq, v['two_ob_sumsq']) and st.count == float(v['two_ob_count'])


@pytest.mark.parametrize('nsr', [False, True])
def test_api_step_fused_equals_call_by_call(eng, oracle_vectors, nsr):
    """es.step's single-synchronisation route leaves exactly what test_params -> rank -> approx_grad -> fit_fn(pheno(0))
    leave: theta, ranker fields, RandomState streams, generation ObStat and the noiseless TrainingResult."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.utils.rankers import CenteredRanker, MultiObjectiveRanker
    from es_pytorch_b200.utils.reporters import ReporterSet
    v = oracle_vectors
    dims, P, table, theta, spec = _small(eng)
    cfg = _Cfg(general=_Cfg(policies_per_gen=12, batch_size=500), policy=_Cfg(l2coeff=0.005))
    comm = dist.world()
    archive = np.random.RandomState(17).randn(16, 2) if nsr else None
    out = []
    for fused in (True, False):
        env, net, policy, nt = _api_objects(eng, table, theta, spec, (64, 64))
        net.set_ob_mean_std(v['obmean'], v['obstd'])
        streams = [np.random.RandomState(1000), np.random.RandomState(1001)]
        fit_fn = BatchedRollout(env, spec.T, coins_per_eval=1, save_obs_chance=0.5, rank_streams=streams, archive=archive)
        ranker = MultiObjectiveRanker(CenteredRanker(), 0.5) if nsr else CenteredRanker()
        assert es._can_fuse_step(comm, policy, fit_fn, ranker)
        for g in range(2):
            if fused:
                tr, ob = es.step(cfg, comm, policy, nt, env, fit_fn, streams[0], ranker, ReporterSet())
            else:
                ob = ObStat(env.observation_space.shape, 0)
                pos, neg, inds, steps = es.test_params(comm, 6, policy, nt, ob, fit_fn, streams[0])
                ranker.rank(pos, neg, inds)
                es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
                tr = fit_fn(policy.pheno(np.zeros(len(policy))), False)
        out.append(dict(theta=policy.flat_params.copy(), w=np.asarray(ranker.ranked_fits).copy(), fits=ranker.fits.copy(),
                        inds=np.asarray(ranker.noise_inds).copy(), n=ranker.n_fits_ranked, res=list(tr.result),
                        steps=tr.steps, ob=(ob.sum.copy(), ob.sumsq.copy(), ob.count),
                        rs=[s.get_state()[1].copy() for s in streams],
                        module=torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy().copy()))
    a, b = out
    assert np.array_equal(a['theta'], b['theta']) and np.array_equal(a['module'], b['module'])
    assert np.array_equal(a['module'], a['theta'])
    assert np.array_equal(a['w'], b['w']) and a['w'].dtype == b['w'].dtype and a['n'] == b['n']
    assert np.array_equal(a['fits'], b['fits']) and np.array_equal(a['inds'], b['inds'])
    assert a['res'] == b['res'] and a['steps'] == b['steps']
    assert np.array_equal(a['ob'][0], b['ob'][0]) and np.array_equal(a['ob'][1], b['ob'][1]) and a['ob'][2] == b['ob'][2] > 0
    assert all(np.array_equal(x, y) for x, y in zip(a['rs'], b['rs']))


def test_api_opaque_fit_fn_loop_matches_oracle(eng):
    """The per-perturbation compatibility path (an opaque python fit_fn like simple_example.py:37-40)."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym import gym_runner
    from es_pytorch_b200.gym.training_result import RewardResult
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.utils.rankers import CenteredRanker
    dims, P, table, theta, spec = _small(eng)
    env, net, policy, nt = _api_objects(eng, table, theta, spec, (64, 64))
    rs = np.random.RandomState(1000)

    def r_fn(model):
        save_obs = rs.random() < 0.0
        rews, behv, obs, steps = gym_runner.run_model(model, env, 10000, rs)
        return RewardResult(rews, behv, obs if save_obs else np.array([np.zeros(env.observation_space.shape)]), steps)

    gen_obstat = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(dist.world(), 5, policy, nt, gen_obstat, r_fn, rs)
    opos, oneg, oinds, osteps, _ = orc.es_test_params(table, theta, 0.02, dims, spec, [1000], 5, np.zeros(17),
                                                      np.ones(17), 5.0, 10000, coins_per_eval=1)
    assert np.array_equal(inds, oinds) and steps == osteps
    assert np.abs(pos - opos).max() <= 4e-4 and np.abs(neg - oneg).max() <= 4e-4
    ranker = CenteredRanker()
    ranker.rank(pos, neg, inds)
    flat, opt = theta.copy(), orc.AdamOracle(P, 0.01)
    w, n = orc.centered_ranker(opos, oneg)
    orc.approx_grad(flat, opt, w, oinds, n, table, 500, 0.005)
    es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
    assert np.array_equal(ranker.ranked_fits, w)
    assert np.abs(policy.flat_params - flat).max() <= 2e-6


# ---- size-independent properties at BASELINE.json's full sizes ---------------------------------------------
def test_full_size_properties(eng):
    """Humanoid-shaped P=29393, K=10000 (config 3): properties that need no CPU replay."""
    g = torch.Generator(device=eng.device).manual_seed(1)
    P, K, L = 29393, 10000, 50_000_000
    table = torch.randn(L, generator=g, device=eng.device, dtype=torch.float32)
    idx = torch.randint(0, L - P, (K,), generator=g, device=eng.device, dtype=torch.int64)
    fpos = torch.randn(K, generator=g, device=eng.device, dtype=torch.float64)
    fneg = torch.randn(K, generator=g, device=eng.device, dtype=torch.float64)
    w, ranks = eng.centered_rank(fpos, fneg, want_ranks=True)
    r = ranks.view(-1).to(torch.int64)
    assert int(r.sum().item()) == (2 * K) * (2 * K - 1) // 2 and int(r.min().item()) == 0       # a permutation
    order = torch.argsort(torch.cat((fpos, fneg)))
    assert torch.equal(r[order].cpu(), torch.arange(2 * K))                                         # sortedness
    w_swapped = eng.centered_rank(fneg, fpos)
    assert torch.equal(w_swapped, -w)                                                                # antisymmetry
    # reconstruction is linear in the weights and a one-hot weight returns its slice
    g1 = eng.grad_reconstruct(table, idx, w, P)
    w2 = torch.rand(K, generator=g, device=eng.device, dtype=torch.float32)
    g2 = eng.grad_reconstruct(table, idx, w2, P)
    g12 = eng.grad_reconstruct(table, idx, w + w2, P)
    scale = float(g12.abs().max().item())
    assert float((g12 - (g1 + g2)).abs().max().item()) <= 1e-5 * scale
    onehot = torch.zeros(K, device=eng.device, dtype=torch.float32)
    onehot[K - 1] = 1.0
    sl = eng.grad_reconstruct(table, idx, onehot, P)
    assert torch.equal(sl, table[int(idx[K - 1].item()):int(idx[K - 1].item()) + P])
    # float64 reference of the same sum on the device, chunked (plumbing only; not the product path)
    ref = torch.zeros(P, dtype=torch.float64, device=eng.device)
    ar = torch.arange(P, device=eng.device)
    for c in range(0, K, 500):
        rows = table[(idx[c:c + 500, None] + ar[None, :])].double()
        ref += (w[c:c + 500, None].double() * rows).sum(0)
    assert float((g1.double() - ref).abs().max().item()) <= 1e-5 * float(ref.abs().max().item())


def test_config3_parity_of_the_tensor_core_modes_vs_float32(eng):
    """BASELINE config 3 (Humanoid-shaped 376-64-64-17, K = 10 000, T = 1000, sigma 0.02): the tensor-core rollouts against the
    float32 CUDA-core rollout on IDENTICAL inputs -- how many of the 20 000 integer ranks differ, the largest change of a rank
    weight and ||g_tc - g_f32|| / ||g_f32|| of the reconstructed gradient (generation.parity_report; bench.py prints the same
    numbers as ``also.parity``).  Centered ranks are a step function of the fitness, so fitness errors at float32 rounding
    level (1e-6 of the spread: what ANY float32 implementation with another summation order has) already move near-tied
    neighbours by one rank; the bounds below are what the split path must stay inside (measured values in DESIGN.md)."""
    from es_pytorch_b200 import _lib
    from es_pytorch_b200.generation import DeviceGeneration, parity_report
    from es_pytorch_b200.nn.optimizers import Adam
    spec = orc.SyntheticEnvSpec(376, 17, 1000)
    sizes = [376, 64, 64, 17]
    P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
    g = torch.Generator(device=eng.device).manual_seed(123)
    table = torch.randn(60_000_000, generator=g, device=eng.device, dtype=torch.float32)
    theta = (np.random.RandomState(7).randn(P) * 0.1).astype(np.float32)
    gen = DeviceGeneration(table, eng.to_device(theta), sizes, eng.to_device(spec.obs_stream), eng.to_device(spec.rew_vec),
                           [np.random.RandomState(1000 + r) for r in range(8)], 0.02, 0.005, Adam(P, 0.01), coins_per_eval=1,
                           rollout_mode=_lib.ES_ROLLOUT_TC3, engine=eng)
    gen.evaluate(1250)                                                   # draws this generation's 10 000 indices
    rep3 = parity_report(gen, _lib.ES_ROLLOUT_TC3, _lib.ES_ROLLOUT_F32)
    rep1 = parity_report(gen, _lib.ES_ROLLOUT_TC, _lib.ES_ROLLOUT_F32)
    print('\nconfig-3 parity tc3 vs f32:', rep3, '\nconfig-3 parity tc  vs f32:', rep1)
    assert rep3['ranks_total'] == 20000
    assert rep3['fitness_rms_err_over_spread'] <= 6e-6
    assert rep3['max_rank_shift'] <= 3 and rep3['max_abs_dw'] <= 3.5 / 19999
    assert rep3['grad_rel_err'] <= 5e-4
    assert rep1['fitness_rms_err_over_spread'] <= 5e-3 and rep1['grad_rel_err'] <= 5e-2
    assert rep3['grad_rel_err'] * 20 <= rep1['grad_rel_err']


def test_api_step_matches_the_real_reference_step_and_its_checkpoint(eng):
    """es.step through the device API against tests/golden/ref_step.npz (the reference's own es.step, three generations):
    after every step the caller's RandomState is where the reference left it -- including the rs.random() of the noiseless
    evaluation (es.py:48) --, theta within 3e-6, the noiseless result within float32 tolerance.  The third step starts from
    tests/golden/policy-ref, the reference's own Policy.save pickle, loaded through the compat shims (Adam m, v, t)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import os, sys
import numpy as np, torch
root = %r
from src.core import es
from src.core.noisetable import NoiseTable
from src.core.policy import Policy
from src.nn.nn import FeedForward
from src.nn.optimizers import Adam
from src.utils.rankers import CenteredRanker
from src.utils.reporters import ReporterSet
from es_pytorch_b200 import dist
from es_pytorch_b200.gym.batched import BatchedRollout
from es_pytorch_b200.gym.synthetic_env import SyntheticEnv
v = np.load(os.path.join(root, 'tests', 'golden', 'ref_step.npz'))
obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
env = SyntheticEnv(obs_dim, act_dim, T)
net = FeedForward([64, 64], torch.nn.Tanh(), env, 0.0, 5)
policy = Policy(net, 0.02, Adam(len(v['theta0']), 0.01))
policy.flat_params[...] = v['theta0']
nt = NoiseTable(len(policy), table)
rs = np.random.RandomState(int(v['seed']))
class Cfg(dict):
    __getattr__ = dict.__getitem__
cfg = Cfg(general=Cfg(policies_per_gen=2 * n_pairs, batch_size=500), policy=Cfg(l2coeff=0.005))
fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']))
comm = dist.world()
for g in range(3):
    if g == 2:
        policy = Policy.load(os.path.join(root, 'tests', 'golden', 'policy-ref'))
        assert policy.optim.t == 2 and np.array_equal(policy.flat_params, v['s1_theta'])
    ranker = CenteredRanker()
    tr, gen_obstat = es.step(cfg, comm, policy, nt, env, fit_fn, rs, ranker, ReporterSet())
    policy.update_obstat(gen_obstat)
    st = rs.get_state()
    assert np.array_equal(st[1], v['s%%d_rs_key' %% g]) and st[2] == int(v['s%%d_rs_pos' %% g]), 'stream after step %%d' %% g
    if g < 2:
        assert np.array_equal(np.asarray(ranker.noise_inds), v['s%%d_inds' %% g])
        assert np.array_equal(gen_obstat.sum, v['s%%d_ob_sum' %% g]) and gen_obstat.count == float(v['s%%d_ob_count' %% g])
    assert np.abs(policy.flat_params - v['s%%d_theta' %% g]).max() <= 3e-6, np.abs(policy.flat_params - v['s%%d_theta' %% g]).max()
    assert abs(tr.result[0] - float(v['s%%d_noiseless' %% g][0])) <= 1e-5 * max(1.0, abs(float(v['s%%d_noiseless' %% g][0]))) * T ** 0.5
print('STEP_OK')
''' % root
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.path.join(root, 'es_pytorch_b200', 'compat'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and 'STEP_OK' in out.stdout, (out.stdout + out.stderr)[-3000:]


@pytest.mark.parametrize('mode', [0, 2])
def test_api_action_noise_generation_matches_the_real_reference(eng, mode):
    """ac_std = 0.01 (configs/simple_conf.json:14, obj.json:18, nsra.json:17) through the reference-facing API, against the
    REAL reference's test_params -> rank -> approx_grad with a noisy FeedForward (tests/golden/ref_step.npz, `acn_*`,
    generated by make_ref_step.py): FeedForward.forward draws rs.randn(act) at every step from the stream that also draws
    the indices and the coins (nn.py:47-48).  On the device that is es_draw_noisy + es_rollout_openloop_noisy: indices,
    obs statistics and the caller's RandomState afterwards (key, position, has_gauss) are exact, the cached gaussian to an
    ulp, fitness to float32 tolerance, rank weights equal, theta within 3e-6."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.nn.obstat import ObStat
    from es_pytorch_b200.utils.rankers import CenteredRanker
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_step.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    env, net, policy, nt = _api_objects(eng, table, v['theta0'].copy(), spec, tuple(int(h) for h in v['hidden']))
    net._action_std = float(v['acn_std'])
    rs = np.random.RandomState(int(v['acn_seed']))
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']), rollout_mode=mode)
    gen_obstat = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(dist.world(), n_pairs, policy, nt, gen_obstat, fit_fn, rs)
    assert np.array_equal(inds, v['acn_inds'])
    st = rs.get_state()
    assert np.array_equal(st[1], v['acn_rs_key']) and st[2] == int(v['acn_rs_pos']) and st[3] == int(v['acn_rs_has_gauss'])
    assert abs(st[4] - float(v['acn_rs_gauss'])) <= 2 * np.spacing(abs(float(v['acn_rs_gauss'])))
    tol = 2e-5 if mode == 0 else 6e-5
    assert np.abs(pos - v['acn_pos']).max() <= tol and np.abs(neg - v['acn_neg']).max() <= tol
    assert np.array_equal(gen_obstat.sum, v['acn_ob_sum']) and gen_obstat.count == float(v['acn_ob_count'])
    ranker = CenteredRanker()
    ranker.rank(pos, neg, inds)
    assert np.array_equal(np.asarray(ranker.ranked_fits).ravel(), v['acn_w'].ravel())
    es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
    assert np.abs(policy.flat_params - v['acn_theta']).max() <= 3e-6


def test_run_model_with_action_noise_stays_on_the_device(eng):
    """The per-policy route (an opaque fit_fn calling gym_runner.run_model with the rank's stream, simple_example.py:37-40)
    with ac_std != 0: the episode is still one fused launch -- the T x act gaussians are drawn from the caller's
    RandomState in one call (same stream consumption as T calls of rs.randn(act)) -- and equals the oracle's step loop."""
    from es_pytorch_b200.gym.gym_runner import run_model
    obs_dim, act_dim, T = 17, 6, 60
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    dims = orc.layer_dims(obs_dim, (64, 64), act_dim)
    theta = (np.random.RandomState(2).randn(orc.n_params(dims)) * 0.1).astype(np.float32)
    table = np.zeros(orc.n_params(dims) + 8, dtype=np.float32)
    env, net, policy, nt = _api_objects(eng, table, theta, spec, (64, 64))
    net._action_std = 0.01
    a, b = np.random.RandomState(77), np.random.RandomState(77)
    a.randn(1)
    b.randn(1)                                              # both start with a cached gaussian
    launches = eng.launches
    rews, behv, obs, step = run_model(policy.pheno(np.zeros(len(policy))), env, T, a)
    assert eng.launches - launches <= 4, 'one normalise + rollout launches, not a step loop'
    rr, bb, _, st = orc.run_model(spec, orc.unflatten(theta, dims), np.zeros(obs_dim), np.ones(obs_dim), 5.0, T, batched=False,
                                  ac_std=0.01, rs=b)
    assert step == st and abs(sum(rews) - sum(rr)) <= 1e-5 * max(1.0, np.abs(rr).sum())
    assert np.allclose(behv[-3:], bb[-3:], rtol=1e-4, atol=1e-5)
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]


@pytest.mark.parametrize('n_streams', [1, 3])
def test_api_step_fused_with_action_noise_matches_the_oracle(eng, n_streams):
    """es.step on the single-synchronisation route with a BatchedRollout and ac_std = 0.01, two generations with the ac_std /
    sigma decays of obj.py:81-82 in between: per generation the device draws indices, coins and T x act gaussians per
    rollout in stream order (es_draw_noisy), the noiseless evaluation draws its coin but no noise (fit_fn(model, False),
    es.py:48).  Against the oracle's es_step: indices and the callers' RandomState objects (key, position, has_gauss)
    exact, cached gaussian to 2 ulp, fitness and theta to float32 tolerance."""
    from es_pytorch_b200 import dist
    from es_pytorch_b200.core import es
    from es_pytorch_b200.gym.batched import BatchedRollout
    from es_pytorch_b200.utils.rankers import CenteredRanker
    from es_pytorch_b200.utils.reporters import Reporter
    obs_dim, act_dim, hidden, T, n = 17, 5, (64, 64), 37, 4          # T * act odd: the gaussian cache crosses rollouts
    spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    dims = orc.layer_dims(obs_dim, hidden, act_dim)
    P = orc.n_params(dims)
    rs0 = np.random.RandomState(5)
    table, theta = rs0.randn(P + 150_000).astype(np.float32), (rs0.randn(P) * 0.1).astype(np.float32)
    env, net, policy, nt = _api_objects(eng, table, theta.copy(), spec, hidden)
    net._action_std = 0.01
    seeds = [910 + 3 * r for r in range(n_streams)]
    streams = [np.random.RandomState(s) for s in seeds]
    ref_streams = [np.random.RandomState(s) for s in seeds]
    streams[0].randn(1); ref_streams[0].randn(1)                      # the first stream starts with a cached gaussian
    fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=0.25, rank_streams=streams)
    cfg = _Cfg(general=_Cfg(policies_per_gen=2 * n, batch_size=500), policy=_Cfg(l2coeff=0.005))
    ranker = CenteredRanker()
    assert es._can_fuse_step(dist.world(), policy, fit_fn, ranker)
    flat, opt = theta.copy(), orc.AdamOracle(P, 0.01)
    stat = orc.ObStatOracle((obs_dim,), 1e-2)
    obmean, obstd, std, ac_std = np.zeros(obs_dim), np.ones(obs_dim), 0.02, 0.01
    for g in range(2):
        tr, gen_obstat = es.step(cfg, dist.world(), policy, nt, env, fit_fn, streams[0], ranker, Reporter())
        policy.update_obstat(gen_obstat)
        ref = orc.es_step(table, flat, opt, std, dims, spec, ref_streams, n, obmean, obstd, 5.0, T, 500, 0.005, coins_per_eval=1,
                          save_obs_chance=0.25, batched=False, ac_std=ac_std)
        stat.inc(ref['obstat'].sum, ref['obstat'].sumsq, ref['obstat'].count)
        obmean, obstd = stat.mean, stat.std
        assert np.array_equal(np.asarray(ranker.noise_inds), ref['inds'])
        assert np.abs(ranker.fits_pos - ref['pos']).max() <= 1e-4 and np.abs(ranker.fits_neg - ref['neg']).max() <= 1e-4
        assert np.array_equal(gen_obstat.sum, ref['obstat'].sum) and gen_obstat.count == ref['obstat'].count
        assert np.abs(policy.flat_params - flat).max() <= 3e-6
        assert abs(tr.result[0] - ref['noiseless'][0]) <= 1e-4
        for a, b in zip(streams, ref_streams):
            sa, sb = a.get_state(), b.get_state()
            assert np.array_equal(sa[1], sb[1]) and sa[2] == sb[2] and sa[3] == sb[3], f'stream state after generation {g}'
            assert abs(sa[4] - sb[4]) <= 2 * np.spacing(abs(sb[4]))
        ac_std *= 0.5; net._action_std = ac_std                       # obj.py:81
        std *= 0.9; policy.std = std                                  # obj.py:82
