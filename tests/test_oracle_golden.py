"""CPU: the oracle against the reference's own known-answer tests and the golden vectors
produced by the real reference modules (tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest

from oracle import es_oracle as orc


# ---- reference known-answer tests, restated against the oracle ------------------------------------
def test_batch_noise_layout():
    """test/utils/utils_test.py:7-21 (stale 3-arg call fixed by inserting policy_len)."""
    params, table = 50, np.arange(100)
    inds = np.arange(40)
    expected = np.array([[i + j for j in range(params)] for i in range(len(inds))])
    assert (next(orc.batch_noise(inds, table, params, len(inds))) == expected).all()
    b = orc.batch_noise(inds, table, params, 19)
    assert (next(b) == expected[:19]).all() and (next(b) == expected[19:38]).all() and (next(b) == expected[38:]).all()


def test_scale_noise_known_answer():
    """test/utils/utils_test.py:24-40: exact equality with the dense dot, batch 3 and full."""
    evals, params = 100, 500
    fits, inds, table = np.arange(evals), np.arange(evals), np.arange(2000)
    expected = np.dot(fits, [[i + j for j in range(params)] for i in range(evals)])
    assert (orc.scale_noise(fits, inds, table, params, 3) == expected).all()
    assert (orc.scale_noise(fits, inds, table, params, evals) == expected).all()
    # the same data as float32 (what the CUDA kernel sees) is still exact: all partial sums < 2^24
    f32 = orc.scale_noise(fits.astype(np.float32), inds, table.astype(np.float32), params, 3)
    assert f32.dtype == np.float32 and (f32 == expected).all()


def test_moo_weighted_rank():
    """test/utils/rankers.py:6-27: MOO rank = per-column centered rank blended by w."""
    x = np.reshape(np.arange(20), (-1, 2)).astype(np.float64)
    pos, neg = x[:5], x[5:]
    for w in (0.5, 0.1):
        got, n = orc.moo_ranker(pos, neg, w)
        r0, r1 = orc.centered_rank(x[:, 0]), orc.centered_rank(x[:, 1])
        y = r0 * np.float32(w) + r1 * np.float32(1 - w)
        assert (got == y[:5] - y[5:]).all() and n == 10


def test_share_results_layout():
    """test/es/es_runner_test.py:10-31 for 3 virtual ranks."""
    evals, objectives, size = 5, 4, 3
    per_rank = []
    for rank in range(size):
        pf = evals * rank + 1
        inds = (np.arange(evals) + pf) * 10
        fp = [[i + i * 10 ** j if j != 0 else i for j in range(objectives)] for i in range(pf, pf + evals)]
        fn = (-np.array(fp)).tolist()
        per_rank.append(np.array([p + n + [i] for p, n, i in zip(fp, fn, inds)], dtype=np.float64))
    res = orc.share_results(per_rank)
    expected = []
    for i in range(1, evals * size + 1):
        p = [i + i * 10 ** j if j != 0 else i for j in range(objectives)]
        expected.append(p + (-np.array(p)).tolist() + [i * 10])
    assert (res == expected).all()


def test_novelty_known_answer():
    """test/utils/novelty_test.py:27-33."""
    beh, archive = np.array([0, 0]), np.array([[2, 2], [1, 1], [3, 3]])
    assert orc.novelty(beh, archive, 1) == np.sqrt(2)
    assert orc.novelty(beh, archive, 2) == (np.sqrt(2) + np.sqrt(8)) / 2
    assert orc.novelty(beh, archive, 3) == (np.sqrt(2) + np.sqrt(8) + np.sqrt(18)) / 3
    assert orc.novelty(beh, archive, 50) == (np.sqrt(2) + np.sqrt(8) + np.sqrt(18)) / 3


def test_obstat_merge():
    """test/utils/obstat_test.py:8-23 for 3 virtual ranks."""
    size, ob = 3, 5
    total = orc.ObStatOracle(ob, 0)
    es, eq = np.zeros(ob), np.zeros(ob)
    for r in range(size):
        o = orc.ObStatOracle(ob, 0)
        o.inc(np.arange(ob) * (r + 1), np.square(np.arange(ob) * (r + 1)), 1)
        total.merge(o)
        es += np.arange(ob) * (r + 1)
        eq += np.square(np.arange(ob) * (r + 1))
    assert (total.sum == es).all() and (total.sumsq == eq).all() and total.count == size


def test_table_content():
    """test/es/noisetable_test.py:26."""
    assert np.isclose(orc.make_noise(5, 1), np.random.RandomState(1).randn(5).astype(np.float32)).all()


# ---- golden vectors from the real reference modules -------------------------------------------------
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_rankers_match_reference(ref_vectors, tag):
    v = ref_vectors
    w, n = orc.centered_ranker(v[f'rank1_{tag}_pos'], v[f'rank1_{tag}_neg'])
    assert w.dtype == np.float32 and np.array_equal(w, v[f'rank1_{tag}_w']) and n == int(v[f'rank1_{tag}_n'])
    for wtag, wt in (('w03', 0.3), ('w10', 1.0), ('w00', 0.0)):
        w2, _ = orc.moo_ranker(v[f'rank2_{tag}_pos'], v[f'rank2_{tag}_neg'], wt)
        assert np.array_equal(w2, v[f'rank2_{tag}_{wtag}_w'])


def test_rank_ties_stable(ref_vectors):
    """Ties: the reference's ``argsort()`` is an unstable sort (the golden vector from the real module shows
    a different order among equal keys than kind='stable' even for n=10), so tie order is unpinned; the
    oracle and the CUDA kernel define it as stable-by-position.  What IS pinned: tied values share the same
    set of ranks, and untied values get identical ranks."""
    x = np.concatenate((ref_vectors['rank_ties_pos'], ref_vectors['rank_ties_neg'])).ravel()
    r = orc.rank(x)
    assert sorted(r.tolist()) == list(range(len(x)))
    for i in range(len(x)):
        for j in range(len(x)):
            if x[i] < x[j]:
                assert r[i] < r[j]
            if x[i] == x[j] and i < j:
                assert r[i] < r[j]          # stable


def test_optimizers_match_reference(ref_vectors):
    v = ref_vectors
    sgd, adam, ses = orc.SGDOracle(40, 0.01), orc.AdamOracle(40, 0.01), orc.SimpleESOracle(40, 0.01)
    for i, g in enumerate(v['sgd_g']):
        assert np.array_equal(sgd.step(g), v['sgd_steps'][i])          # float32 module: bit-exact
        assert np.array_equal(ses.step(g), v['simple_steps'][i])
        a = adam.step(g)
        assert a.dtype == np.float32
        # the real Adam runs in float64 under numpy 2 (np.float64 scalar `a`); float32 pinning differs by rounding only
        assert np.allclose(a, v['adam_steps_real_f64'][i], rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_mt_restatement_matches_numpy(ref_vectors, tag):
    v = ref_vectors
    seed, n, ub, extra = [int(x) for x in v[f'mt_{tag}_cfg']]
    idx, ext, key, pos = orc.mt_draw_indices(v[f'mt_{tag}_key0'], int(v[f'mt_{tag}_pos0']), n, ub, extra)
    assert idx == v[f'mt_{tag}_idx'].tolist()
    assert np.array_equal(np.array(ext, dtype=np.uint32).reshape(n, extra), v[f'mt_{tag}_extra'])
    assert np.array_equal(np.array(key, dtype=np.uint32), v[f'mt_{tag}_key1']) and pos == int(v[f'mt_{tag}_pos1'])


def test_coin_words_are_random_sample():
    rs = np.random.RandomState(3)
    st = rs.get_state()
    u = rs.random()
    r2 = np.random.RandomState()
    r2.set_state(st)
    a, b = (int.from_bytes(r2.bytes(4), 'little') for _ in range(2))
    assert orc.words_to_double(a, b) == u


# ---- frozen oracle outputs (unpinned parts) ------------------------------------------------------------
def test_oracle_frozen_vectors(oracle_vectors):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden import small_problem
    v = oracle_vectors
    dims, P, table, theta, env = small_problem()
    noise = orc.table_get(table, 12345, P)
    assert np.array_equal(orc.pheno_params(theta, 0.02, noise), v['pheno_pos'])
    assert np.array_equal(orc.pheno_params(theta, 0.02, -noise), v['pheno_neg'])
    layers = orc.unflatten(v['pheno_pos'], dims)
    rews, behv, obs, step = orc.run_model(env, layers, v['obmean'], v['obstd'], 5.0, env.T, batched=False)
    assert np.allclose(rews, v['rollout_rews'], rtol=1e-5, atol=1e-6) and step == int(v['rollout_step'])
    # the batched evaluation agrees with the per-step loop to float32 rounding
    rews_b, _, _, _ = orc.run_model(env, layers, v['obmean'], v['obstd'], 5.0, env.T, batched=True)
    assert np.allclose(rews_b, rews, rtol=1e-4, atol=1e-5)
    assert np.array_equal(orc.normalise_obs(env.obs_stream[:env.T], v['obmean'], v['obstd'], 5.0), v['obsn'])


# ---- the rankers outside the north-star pair (rankers.py:61-103), pinned by the real module ----------------------
RK = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_rankers.npz'))
RK_TAGS = ('a', 'b', 'c', 'd')


@pytest.mark.parametrize('tag', RK_TAGS)
@pytest.mark.parametrize('name', ['double_positive', 'max_normalized', 'semi_centered'])
def test_shaped_rankers_match_reference(tag, name):
    w, n = orc.shaped_ranker(RK[f'{tag}_pos'], RK[f'{tag}_neg'], name)
    ref = RK[f'{tag}_{name}_w']
    assert w.dtype == ref.dtype and w.shape == ref.shape and n == int(RK[f'{tag}_{name}_n'])
    assert np.array_equal(w, ref)                                   # bit-exact, float32 and float64 shapings alike
    w2, _ = orc.shaped_ranker(RK[f'{tag}_pos2'], RK[f'{tag}_neg2'], name, w=0.3)
    ref2 = RK[f'{tag}_moo_{name}_w']
    assert w2.dtype == ref2.dtype and np.array_equal(w2.reshape(ref2.shape), ref2)


@pytest.mark.parametrize('tag', RK_TAGS)
@pytest.mark.parametrize('name', ['centered', 'double_positive', 'max_normalized'])
@pytest.mark.parametrize('ptag,pct', [('p00', 0.0), ('p10', 0.1), ('p50', 0.5), ('p100', 1.0)])
def test_elite_ranker_matches_reference(tag, name, ptag, pct):
    vals, inds, fit, n = orc.elite_ranker(RK[f'{tag}_pos'], RK[f'{tag}_neg'], RK[f'{tag}_inds'], name, pct)
    assert n == int(RK[f'{tag}_elite_{name}_{ptag}_n']) == len(vals)
    order = np.lexsort((inds, vals))                                # the reference's order is unspecified: compare sets
    ref_v = RK[f'{tag}_elite_{name}_{ptag}_vals']
    assert vals.dtype == ref_v.dtype
    assert np.array_equal(vals[order], ref_v)
    assert np.array_equal(inds[order], RK[f'{tag}_elite_{name}_{ptag}_inds'])


# ---- the REAL reference pipeline, run in the build container by tests/golden/make_ref_pipeline.py --------------------
def test_oracle_reproduces_the_real_reference_pipeline():
    """Two generations of src.core.es.test_params -> CenteredRanker.rank -> es.approx_grad -> Policy.update_obstat executed
    by the reference's own code (real Policy.pheno / FeedForward.forward / gym_runner.run_model / RewardResult / ObStat / Adam).
    Pins what no reference test pins: the RNG interleaving, pheno, the forward, run_model, approx_grad, the obs feedback."""
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    dims = orc.layer_dims(obs_dim, tuple(int(h) for h in v['hidden']), act_dim)
    P = orc.n_params(dims)
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    env = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    flat = v['theta0'].copy()
    assert len(flat) == P
    opt = orc.AdamOracle(P, 0.01)
    rs = np.random.RandomState(int(v['seed']))
    stat = orc.ObStatOracle((obs_dim,), 1e-2)                               # Policy.obstat, policy.py:27
    obmean, obstd = np.zeros(obs_dim), np.ones(obs_dim)
    for g in range(2):
        assert np.array_equal(obmean, v[f'g{g}_obmean']) and np.array_equal(obstd, v[f'g{g}_obstd'])
        pos, neg, inds, steps, gen_stat = orc.es_test_params(table, flat, 0.02, dims, env, [0], n_pairs, obmean, obstd, 5.0, T,
                                                            coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']),
                                                            batched=False, rank_states=[rs])
        assert np.array_equal(inds, v[f'g{g}_inds']) and steps == int(v[f'g{g}_steps'])        # RNG interleaving: exact
        assert pos.dtype == v[f'g{g}_pos'].dtype and pos.shape == v[f'g{g}_pos'].shape
        # pheno, forward, run_model: bit-exact when torch's CPU matmul takes the same path as in the generating run; its
        # blocking depends on the thread count other tests may have set, hence a float32-ulp tolerance
        scale = max(1.0, float(np.abs(v[f'g{g}_pos']).max()))
        assert np.abs(pos - v[f'g{g}_pos']).max() <= 1e-6 * scale and np.abs(neg - v[f'g{g}_neg']).max() <= 1e-6 * scale
        assert np.array_equal(gen_stat.sum, v[f'g{g}_ob_sum']) and np.array_equal(gen_stat.sumsq, v[f'g{g}_ob_sumsq'])
        assert gen_stat.count == float(v[f'g{g}_ob_count']) > 0
        stat.inc(gen_stat.sum, gen_stat.sumsq, gen_stat.count)              # Policy.update_obstat, policy.py:69-71
        obmean, obstd = stat.mean, stat.std
        w, n_ranked = orc.centered_ranker(pos, neg)
        assert np.array_equal(w, v[f'g{g}_w']) and n_ranked == int(v[f'g{g}_n_ranked'])
        orc.approx_grad(flat, opt, w, inds, n_ranked, table, 500, 0.005)
        # the real Adam computes its step in float64 under numpy 2 (float32 under the reference's numpy 1.18)
        assert np.abs(flat - v[f'g{g}_theta']).max() <= 2e-6
    assert np.array_equal(rs.get_state()[1], v['rs_key']) and rs.get_state()[2] == int(v['rs_pos'])
    rs.random()                                                             # the noiseless evaluation's coin
    layers = orc.unflatten(orc.pheno_params(flat, 0.02, None), dims)
    rews, behv, _, _ = orc.run_model(env, layers, obmean, obstd, 5.0, T, batched=False)
    assert abs(orc.reward_result(rews)[0] - float(v['noiseless_result'][0])) <= 1e-5
    assert np.allclose(behv[-3:-1], v['noiseless_behv'], rtol=1e-5, atol=1e-6)


def _ref_pipeline_setup():
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_pipeline.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    dims = orc.layer_dims(obs_dim, tuple(int(h) for h in v['hidden']), act_dim)
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    return v, obs_dim, act_dim, T, n_pairs, dims, table, orc.SyntheticEnvSpec(obs_dim, act_dim, T)


def test_oracle_reproduces_the_real_reference_nsra_generation():
    """Real NSRResult (reward + novelty of the final (x, y) vs an archive, k = 10) through the real test_params, real
    MultiObjectiveRanker(CenteredRanker(), 0.5) and approx_grad."""
    v, obs_dim, act_dim, T, n_pairs, dims, table, env = _ref_pipeline_setup()
    flat, opt = v['theta0'].copy(), orc.AdamOracle(len(v['theta0']), 0.01)
    rs = np.random.RandomState(int(v['nsra_seed']))
    pos, neg, inds, _, _ = orc.es_test_params(table, flat, 0.02, dims, env, [0], n_pairs, np.zeros(obs_dim), np.ones(obs_dim), 5.0,
                                              T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']),
                                              archive=v['nsra_archive'], nov_k=10, batched=False, rank_states=[rs])
    assert np.array_equal(inds, v['nsra_inds']) and pos.shape == v['nsra_pos'].shape == (n_pairs, 2)
    assert np.abs(pos - v['nsra_pos']).max() <= 1e-6 * max(1.0, np.abs(v['nsra_pos']).max())
    assert np.abs(neg - v['nsra_neg']).max() <= 1e-6 * max(1.0, np.abs(v['nsra_neg']).max())
    w, n_ranked = orc.moo_ranker(pos, neg, 0.5)
    assert np.array_equal(w, v['nsra_w'])
    orc.approx_grad(flat, opt, w, inds, n_ranked, table, 500, 0.005)
    assert np.abs(flat - v['nsra_theta']).max() <= 2e-6


def test_oracle_reproduces_the_real_reference_elite_update():
    """obj.py:50's EliteRanker(CenteredRanker(), elite) through the real approx_grad: the elite keep the noise index of their
    pair whatever their sign, nothing is subtracted, the sum is divided by the number of elite."""
    v, obs_dim, act_dim, T, n_pairs, dims, table, env = _ref_pipeline_setup()
    flat, opt = v['theta0'].copy(), orc.AdamOracle(len(v['theta0']), 0.01)
    rs = np.random.RandomState(int(v['elite_seed']))
    pos, neg, inds, _, _ = orc.es_test_params(table, flat, 0.02, dims, env, [0], n_pairs, np.zeros(obs_dim), np.ones(obs_dim), 5.0,
                                              T, coins_per_eval=1, save_obs_chance=0.0, batched=False, rank_states=[rs])
    assert np.array_equal(inds, v['elite_inds'])
    vals, sel, _, n = orc.elite_ranker(pos, neg, inds, 'centered', float(v['elite_pct']))
    order = np.lexsort((sel, vals))
    assert n == int(v['elite_n']) and np.array_equal(vals[order], v['elite_vals']) and np.array_equal(sel[order], v['elite_sel'])
    orc.approx_grad(flat, opt, vals, sel, n, table, 500, 0.005)
    assert np.abs(flat - v['elite_theta']).max() <= 2e-6


def test_oracle_reproduces_the_real_reference_two_rank_layout():
    """The real test_params on two (thread-emulated) MPI ranks, each with its own RandomState: rank-major rows out of
    es._share_results, summed steps, ObStat.mpi_inc -- what a process carrying two 'virtual ranks' must reproduce."""
    v, obs_dim, act_dim, T, _, dims, table, env = _ref_pipeline_setup()
    seeds, n = [int(x) for x in v['two_seeds']], int(v['two_n'])
    pos, neg, inds, steps, st = orc.es_test_params(table, v['theta0'].copy(), 0.02, dims, env, seeds, n, np.zeros(obs_dim),
                                                   np.ones(obs_dim), 5.0, T, coins_per_eval=1,
                                                   save_obs_chance=float(v['save_obs_chance']), batched=False)
    assert np.array_equal(inds, v['two_inds']) and steps == int(v['two_steps'])
    assert np.abs(pos - v['two_pos']).max() <= 1e-6 * max(1.0, np.abs(v['two_pos']).max())
    assert np.abs(neg - v['two_neg']).max() <= 1e-6 * max(1.0, np.abs(v['two_neg']).max())
    assert np.array_equal(st.sum, v['two_ob_sum']) and np.array_equal(st.sumsq, v['two_ob_sumsq']) and st.count == float(v['two_ob_count'])


@pytest.mark.parametrize('tag', ['sgd', 'simple'])
def test_oracle_reproduces_the_real_reference_other_optimizers(tag):
    """Momentum SGD (two consecutive updates) and SimpleES through the real es.approx_grad."""
    v, obs_dim, act_dim, T, n_pairs, dims, table, env = _ref_pipeline_setup()
    P = len(v['theta0'])
    flat = v['theta0'].copy()
    opt = orc.SGDOracle(P, 0.01) if tag == 'sgd' else orc.SimpleESOracle(P, 0.01)
    rs = np.random.RandomState(5000)
    for g in range(2):
        pos, neg, inds, _, _ = orc.es_test_params(table, flat, 0.02, dims, env, [0], n_pairs, np.zeros(obs_dim), np.ones(obs_dim),
                                                  5.0, T, coins_per_eval=1, batched=False, rank_states=[rs])
        assert np.array_equal(inds, v[f'{tag}_g{g}_inds'])
        w, n_ranked = orc.centered_ranker(pos, neg)
        orc.approx_grad(flat, opt, w, inds, n_ranked, table, 500, 0.005)
        assert np.abs(flat - v[f'{tag}_g{g}_theta']).max() <= 2e-6


def test_oracle_reproduces_the_real_reference_humanoid_shape():
    """The bench's policy shape (376-64-64-17) through the real pipeline (short episode): flat parameter layout of a wide
    first layer, rank weights, Adam update."""
    v, _, _, _, _, _, table, _ = _ref_pipeline_setup()
    dims = orc.layer_dims(376, (64, 64), 17)
    env = orc.SyntheticEnvSpec(376, 17, 16)
    flat, opt = v['hum_theta0'].copy(), orc.AdamOracle(len(v['hum_theta0']), 0.01)
    assert len(flat) == orc.n_params(dims) == 29393
    rs = np.random.RandomState(6000)
    pos, neg, inds, _, _ = orc.es_test_params(table, flat, 0.02, dims, env, [0], 3, np.zeros(376), np.ones(376), 5.0, 16,
                                              coins_per_eval=1, batched=False, rank_states=[rs])
    assert np.array_equal(inds, v['hum_inds'])
    assert np.abs(pos - v['hum_pos']).max() <= 1e-6 * max(1.0, np.abs(v['hum_pos']).max())
    assert np.abs(neg - v['hum_neg']).max() <= 1e-6 * max(1.0, np.abs(v['hum_neg']).max())
    w, n_ranked = orc.centered_ranker(pos, neg)
    assert np.array_equal(w, v['hum_w'])
    orc.approx_grad(flat, opt, w, inds, n_ranked, table, 500, 0.005)
    assert np.abs(flat - v['hum_theta']).max() <= 2e-6


# ---- the REAL reference es.step / Policy.save / action noise, run by tests/golden/make_ref_step.py ---------------------------
def _ref_step_setup():
    v = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_step.npz'))
    obs_dim, act_dim, T, n_pairs = [int(x) for x in v['cfg']]
    dims = orc.layer_dims(obs_dim, tuple(int(h) for h in v['hidden']), act_dim)
    table = np.random.RandomState(int(v['table_seed'])).randn(int(v['table_len'])).astype(np.float32)
    return v, obs_dim, act_dim, T, n_pairs, dims, table, orc.SyntheticEnvSpec(obs_dim, act_dim, T)


def test_oracle_reproduces_the_real_reference_step_including_the_noiseless_coin():
    """Three generations of the reference's own es.step (es.py:23-51) with an obj.py-style fit_fn: the stream position
    after every step includes the rs.random() of the noiseless evaluation; the third step starts from the reference's own
    checkpoint (Adam m, v, t restored by Policy.load)."""
    v, obs_dim, act_dim, T, n_pairs, dims, table, env = _ref_step_setup()
    P = orc.n_params(dims)
    flat, opt = v['theta0'].copy(), orc.AdamOracle(P, 0.01)
    rs = np.random.RandomState(int(v['seed']))
    stat = orc.ObStatOracle((obs_dim,), 1e-2)
    obmean, obstd = np.zeros(obs_dim), np.ones(obs_dim)
    for g in range(3):
        if g == 2:                                              # resumed from the reference's pickle
            assert opt.t == int(v['ckpt_t']) == 2
            assert np.abs(opt.m - v['ckpt_m']).max() <= 1e-7 and np.abs(opt.v - v['ckpt_v']).max() <= 1e-9
        out = orc.es_step(table, flat, opt, 0.02, dims, env, [rs], n_pairs, obmean, obstd, 5.0, T, 500, 0.005, coins_per_eval=1,
                          save_obs_chance=float(v['save_obs_chance']), batched=False)
        if g < 2:
            assert np.array_equal(out['inds'], v[f's{g}_inds'])
            assert np.array_equal(out['obstat'].sum, v[f's{g}_ob_sum']) and out['obstat'].count == float(v[f's{g}_ob_count'])
        st = rs.get_state()
        assert np.array_equal(st[1], v[f's{g}_rs_key']) and st[2] == int(v[f's{g}_rs_pos']), f'stream after step {g}'
        assert np.abs(flat - v[f's{g}_theta']).max() <= 3e-6
        assert abs(out['noiseless'][0] - float(v[f's{g}_noiseless'][0])) <= 1e-5
        stat.inc(out['obstat'].sum, out['obstat'].sumsq, out['obstat'].count)
        obmean, obstd = stat.mean, stat.std
    assert np.array_equal(stat.sum, v['ckpt_obstat_sum'] + out['obstat'].sum)


def test_oracle_reproduces_the_real_reference_action_noise_generation():
    """ac_std = 0.01 (configs/simple_conf.json:14): FeedForward.forward draws rs.randn(act) at every step of every rollout
    from the stream that also draws indices and coins (nn.py:47-48).  Indices and the final stream state (key, position
    AND the cached polar-method gaussian) are exact; fitness to float32 rounding."""
    v, obs_dim, act_dim, T, n_pairs, dims, table, env = _ref_step_setup()
    rs = np.random.RandomState(int(v['acn_seed']))
    flat = v['theta0'].copy()
    pos, neg, inds, steps, stat = orc.es_test_params(table, flat, 0.02, dims, env, [0], n_pairs, np.zeros(obs_dim), np.ones(obs_dim),
                                                     5.0, T, coins_per_eval=1, save_obs_chance=float(v['save_obs_chance']),
                                                     batched=False, rank_states=[rs], ac_std=float(v['acn_std']))
    assert np.array_equal(inds, v['acn_inds'])
    st = rs.get_state()
    assert np.array_equal(st[1], v['acn_rs_key']) and st[2] == int(v['acn_rs_pos'])
    assert st[3] == int(v['acn_rs_has_gauss']) and st[4] == float(v['acn_rs_gauss'])
    assert np.abs(pos - v['acn_pos']).max() <= 2e-6 and np.abs(neg - v['acn_neg']).max() <= 2e-6
    assert np.array_equal(stat.sum, v['acn_ob_sum']) and stat.count == float(v['acn_ob_count'])
    w, n_ranked = orc.centered_ranker(pos, neg)
    assert np.array_equal(w, v['acn_w'])
    orc.approx_grad(flat, orc.AdamOracle(len(flat), 0.01), w, inds, n_ranked, table, 500, 0.005)
    assert np.abs(flat - v['acn_theta']).max() <= 2e-6


def test_reference_checkpoint_unpickles_through_the_shims():
    """tests/golden/policy-ref was written by the reference's own Policy.save (src.core.policy.Policy, src.nn.nn.FeedForward,
    src.nn.optimizers.Adam, src.nn.obstat.ObStat).  It must load here with its Adam moments (the reference keeps m / v as
    plain attributes), and survive a save / load round trip before any optimizer step (no GPU needed for either)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import os, pickle, sys, tempfile
import numpy as np
from src.core.policy import Policy
v = np.load(os.path.join(%r, 'tests', 'golden', 'ref_step.npz'))
p = Policy.load(os.path.join(%r, 'tests', 'golden', 'policy-ref'))
import es_pytorch_b200.core.policy as mine, es_pytorch_b200.nn.optimizers as opt
assert type(p) is mine.Policy and type(p.optim) is opt.Adam and p.optim.t == int(v['ckpt_t']) == 2
assert p.flat_params.dtype == np.float32 and np.array_equal(p.flat_params, v['s1_theta'])
assert np.array_equal(p.optim._host_restore['m'], v['ckpt_m']) and np.array_equal(p.optim._host_restore['v'], v['ckpt_v'])
assert np.array_equal(p.obstat.sum, v['ckpt_obstat_sum']) and p.obstat.count == float(v['ckpt_obstat_count'])
assert p._theta_dev is None and p._module.is_tanh_mlp() and p._module.layer_sizes() == [17, 64, 64, 6]
assert np.array_equal(mine.Policy.get_flat(p._module), v['s1_theta'])          # Policy.load scattered flat_params into the module
d = tempfile.mkdtemp()
p.save(d, 'again')                                                             # before any step: the moments must survive
q = mine.Policy.load(os.path.join(d, 'policy-again'))
assert np.array_equal(q.optim._host_restore['m'], v['ckpt_m']) and np.array_equal(q.optim._host_restore['v'], v['ckpt_v']) and q.optim.t == 2
state = q.optim.__getstate__()
assert set(state) >= {'lr', 'dim', 't', 'beta1', 'beta2', 'epsilon', 'm', 'v'} and '_dev' not in state
print('CKPT_OK')
''' % (root, root)
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.path.join(root, 'es_pytorch_b200', 'compat'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and 'CKPT_OK' in out.stdout, (out.stdout + out.stderr)[-3000:]


def test_closed_loop_oracle_frozen_vectors():
    """The closed-loop variant has no reference implementation: the oracle is its definition, pinned by vectors frozen at
    the time the device kernel was validated against it (tests/golden/make_closed_golden.py)."""
    import importlib.util
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    spec = importlib.util.spec_from_file_location('make_closed_golden', os.path.join(here, 'make_closed_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = np.load(os.path.join(here, 'closed_loop.npz'))
    got = mod.compute()
    assert set(got) == set(want.files)
    for k in want.files:
        a, b = np.asarray(got[k]), want[k]
        # (torch's CPU matrix-vector kernels may differ in the last bit between builds: float32 tolerance, indices exact)
        if k in ('gen_inds', 'gen_steps', 'ob_count', 'env_a', 'env_b'):
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, rtol=0, atol=5e-6), (k, np.abs(a - b).max())
