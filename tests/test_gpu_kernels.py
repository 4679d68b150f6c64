"""GPU parity: every kernel of libes_b200.so, called through the C ABI (ctypes engine),
against the CPU oracle and the committed golden vectors.  Integer / index / rank results
are bit-exact; float results carry their tolerance in the assert."""
import numpy as np
import pytest
import torch

from oracle import es_oracle as orc

pytestmark = pytest.mark.gpu


def dev(eng, a, dtype=None):
    return eng.to_device(np.ascontiguousarray(a), dtype)


# ------------------------------------------------------------------------------------------- a2
@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_draw_indices_golden(eng, ref_vectors, tag):
    v = ref_vectors
    seed, n, ub, extra = [int(x) for x in v[f'mt_{tag}_cfg']]
    key = dev(eng, v[f'mt_{tag}_key0'].astype(np.uint32).view(np.int32).reshape(1, -1))
    pos = dev(eng, np.array([int(v[f'mt_{tag}_pos0'])], dtype=np.int32))
    idx, ext = eng.draw_indices(key, pos, n, ub, extra)
    assert np.array_equal(idx.cpu().numpy(), v[f'mt_{tag}_idx'])
    if extra:
        assert np.array_equal(ext.cpu().numpy().view(np.uint32), v[f'mt_{tag}_extra'])
    assert np.array_equal(key.cpu().numpy().view(np.uint32).ravel(), v[f'mt_{tag}_key1'])
    assert int(pos.item()) == int(v[f'mt_{tag}_pos1'])


@pytest.mark.parametrize('extra', [0, 4])
def test_draw_indices_streams_continue(eng, extra):
    """8 virtual ranks, three consecutive generations: the device streams stay in lock step with numpy."""
    R, n, ub = 8, 333, 250_000_000 - 29393
    streams = [np.random.RandomState(1000 + r) for r in range(R)]
    for s in streams[::2]:
        s.randn(3)                         # arbitrary starting position / cached gaussian
    key = dev(eng, np.stack([s.get_state()[1].astype(np.uint32) for s in streams]).view(np.int32))
    pos = dev(eng, np.array([s.get_state()[2] for s in streams], dtype=np.int32))
    for gen in range(3):
        idx, ext = eng.draw_indices(key, pos, n, ub, extra)
        ref_idx, ref_ext = [], []
        for s in streams:
            for _ in range(n):
                ref_idx.append(int(s.randint(0, ub)))
                ref_ext.append([int.from_bytes(s.bytes(4), 'little') for _ in range(extra)])
        assert np.array_equal(idx.cpu().numpy(), np.array(ref_idx))
        if extra:
            assert np.array_equal(ext.cpu().numpy().view(np.uint32), np.array(ref_ext, dtype=np.uint32))
    assert np.array_equal(key.cpu().numpy().view(np.uint32), np.stack([s.get_state()[1] for s in streams]))
    assert pos.cpu().numpy().tolist() == [s.get_state()[2] for s in streams]


@pytest.mark.parametrize('N,coins', [(7, 1), (96, 1), (3001, 0), (6000, 1)])
def test_draw_noisy_matches_numpy_stream(eng, N, coins):
    """es_draw_noisy against numpy's legacy RandomState in the reference's program order (es.py:66-72 with the fit_fn of
    simple_example.py:37-40 and nn.py:47-48): per pair randint, then per evaluation the coin and N gaussians.  Indices,
    coin words and the stream state (key, position, has_gauss) are bit-exact -- over two consecutive generations, from
    streams that start with and without a cached gaussian, for odd and even N (the cached value crosses evaluations, pairs
    and generations) --; the gaussians agree to the last bit of float32 (CUDA's log vs glibc's: <= 1 ulp of float64)."""
    R, n, ub, scale = 3, 5, 250_000_000 - 29393, 0.01
    streams = [np.random.RandomState(4000 + r) for r in range(R)]
    streams[1].randn(3)                                    # leaves a cached gaussian
    streams[2].randint(0, 1000, size=700)                  # mid-block position
    st = [s.get_state() for s in streams]
    key = dev(eng, np.stack([x[1].astype(np.uint32) for x in st]).view(np.int32))
    pos = dev(eng, np.array([x[2] for x in st], dtype=np.int32))
    has = dev(eng, np.array([x[3] for x in st], dtype=np.int32))
    gau = dev(eng, np.array([x[4] for x in st], dtype=np.float64))
    for gen in range(2):
        idx, coin, noise = eng.draw_noisy(key, pos, has, gau, n, ub, coins, N, scale)
        eng.sync()
        ref_idx, ref_coin, ref_noise = [], [], []
        for s in streams:
            for _ in range(n):
                ref_idx.append(int(s.randint(0, ub)))
                words = []
                for _sgn in range(2):
                    words += [int.from_bytes(s.bytes(4), 'little') for _ in range(2 * coins)]
                    ref_noise.append((s.randn(N) * scale).astype(np.float32))
                ref_coin.append(words)
        assert np.array_equal(idx.cpu().numpy(), np.array(ref_idx))
        if coins:
            assert np.array_equal(coin.cpu().numpy().view(np.uint32), np.array(ref_coin, dtype=np.uint32))
        got, want = noise.cpu().numpy().reshape(-1, N), np.stack(ref_noise)
        assert np.abs(got - want).max() <= np.spacing(np.float32(np.abs(want).max())), np.abs(got - want).max()
        assert (got == want).mean() > 0.9999
        ref_st = [s.get_state() for s in streams]
        assert np.array_equal(key.cpu().numpy().view(np.uint32), np.stack([x[1] for x in ref_st]))
        assert pos.cpu().numpy().tolist() == [x[2] for x in ref_st]
        assert has.cpu().numpy().tolist() == [x[3] for x in ref_st]
        g_dev, g_ref = gau.cpu().numpy(), np.array([x[4] for x in ref_st])
        assert np.all(np.abs(g_dev - g_ref) <= 2 * np.spacing(np.abs(g_ref))), (g_dev, g_ref)


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('obs,act,T,n_pairs', [(17, 6, 150, 9), (24, 9, 130, 160)])
def test_rollout_with_action_noise(eng, mode, obs, act, T, n_pairs):
    """es_rollout_openloop_noisy: the action of every step gets its noise term before reward and position see it
    (nn.py:47-48, gym_runner.py:52-53).  The float32 kernels (general: 9 pairs; packed-FMA: 160 pairs) against the oracle's
    run_model with the same noise values; the tensor-core kernels against the float32 kernel."""
    rs = np.random.RandomState(obs + T + n_pairs)
    dims = orc.layer_dims(obs, (64, 64), act)
    P = orc.n_params(dims)
    L = P + 300_000
    table, theta = rs.randn(L).astype(np.float32), (rs.randn(P) * 0.1).astype(np.float32)
    idx = rs.randint(0, L - P - 1, size=n_pairs).astype(np.int64)
    env = orc.SyntheticEnvSpec(obs, act, T)
    noise = (rs.randn(n_pairs, 2, T, act) * 0.05).astype(np.float32)
    obsn = eng.normalise_obs(dev(eng, env.obs_stream[:T]), dev(eng, np.zeros(obs)), dev(eng, np.ones(obs)), 5.0)

    def run(md, nz):
        fit = torch.zeros(2, n_pairs, dtype=torch.float64, device=eng.device)
        behv = torch.zeros(2, n_pairs, 3, dtype=torch.float32, device=eng.device)
        eng.rollout(dev(eng, table), dev(eng, idx), dev(eng, theta), 0.02, [obs, 64, 64, act], obsn, dev(eng, env.rew_vec),
                    env.pos_scale, fit[0], fit[1], 1, behv[0], behv[1], md, act_noise=None if nz is None else dev(eng, nz))
        eng.sync()
        return fit.cpu().numpy(), behv.cpu().numpy()

    f, b = run(mode, noise)
    f0, _ = run(mode, None)
    assert np.abs(f - f0).max() > 1e-3, 'the noise must change the fitness'
    if mode == 0:
        class _Replay:                                       # run_model draws rs.randn(act) per step: replay the array
            def __init__(self, a): self.a, self.i = a.astype(np.float64), 0
            def randn(self, n): self.i += 1; return self.a[self.i - 1]
        for k in list(range(min(n_pairs, 4))) + [n_pairs - 1]:
            eps = orc.table_get(table, int(idx[k]), P)
            for s, nz in ((0, eps), (1, -eps)):
                layers = orc.unflatten(orc.pheno_params(theta, 0.02, nz), dims)
                rews, bb, _, _ = orc.run_model(env, layers, np.zeros(obs), np.ones(obs), 5.0, T, batched=True, ac_std=1.0,
                                               rs=_Replay(noise[k, s]))
                assert abs(f[s, k] - orc.reward_result(rews)[0]) <= 1e-5 * max(1.0, np.abs(rews).sum())
                assert np.allclose(b[s, k], bb[-3:], rtol=1e-4, atol=1e-5)
    else:
        f32, b32 = run(0, noise)
        spread = max(f32.std(), 1e-3 * np.sqrt(T))
        tol = 6e-6 if mode == 2 else 5e-3
        assert np.sqrt(((f - f32) ** 2).mean()) <= tol * spread + (1e-6 if mode == 2 else 2e-4)
        assert np.abs(b - b32).max() <= (2e-6 if mode == 2 else 2e-3) * 0.05 * T + 1e-4


def test_draw_indices_errors(eng):
    from es_pytorch_b200._lib import EsLibraryError
    key = dev(eng, np.zeros((1, 624), dtype=np.int32))
    pos = dev(eng, np.array([624], dtype=np.int32))
    with pytest.raises(EsLibraryError):           # network larger than the table (noisetable.py:39 ValueError)
        eng.draw_indices(key, pos, 4, 0, 0)
    with pytest.raises(EsLibraryError):
        eng.draw_indices(key, pos, 4, 1 << 33, 0)


# ------------------------------------------------------------------------------------------- a3
def test_perturb_bit_exact(eng):
    rs = np.random.RandomState(0)
    P, L = 5702, 100_000
    table = rs.randn(L).astype(np.float32)
    theta = (rs.randn(P) * 0.1).astype(np.float32)
    idx = np.array([0, 1, 17, L - P - 1, 4242], dtype=np.int64)
    op, on = eng.perturb(dev(eng, theta), dev(eng, table), dev(eng, idx), 0.02)
    for i, k in enumerate(idx):
        noise = orc.table_get(table, int(k), P)
        assert np.array_equal(op[i].cpu().numpy(), orc.pheno_params(theta, 0.02, noise))
        assert np.array_equal(on[i].cpu().numpy(), orc.pheno_params(theta, 0.02, -noise))


# ------------------------------------------------------------------------------------------- a4 / a14
def test_normalise_and_colsum_bit_exact(eng, oracle_vectors):
    rs = np.random.RandomState(1)
    obs = (rs.randn(257, 376) * 3).astype(np.float32)
    mean, std = rs.randn(376) * 0.2, 0.3 + rs.rand(376)
    out = eng.normalise_obs(dev(eng, obs), dev(eng, mean), dev(eng, std), 5.0)
    assert np.array_equal(out.cpu().numpy(), orc.normalise_obs(obs, mean, std, 5.0))
    s, q = eng.obs_colsum(dev(eng, obs))
    es, eq, _ = orc.ob_sum_sq_cnt(obs)
    assert np.array_equal(s.cpu().numpy(), es) and np.array_equal(q.cpu().numpy(), eq)


def test_obstat_accumulate_coins(eng):
    rs = np.random.RandomState(2)
    d, T, n = 17, 40, 500
    s, q = rs.randn(d).astype(np.float32), rs.rand(d).astype(np.float32)
    words = rs.randint(0, 1 << 32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
    chance = 0.07
    osum = dev(eng, np.zeros(d)); osq = dev(eng, np.zeros(d)); cnt = dev(eng, np.zeros(2))
    eng.obstat_accumulate_coins(osum, osq, cnt, dev(eng, s), dev(eng, q), T, dev(eng, words.view(np.int32)), chance)
    ob = orc.ObStatOracle((d,), 0)
    saved = 0
    for a, b in words:
        if orc.words_to_double(int(a), int(b)) < chance:
            ob.inc(s, q, T)
            saved += 1
    assert saved > 5
    assert np.array_equal(osum.cpu().numpy(), ob.sum) and np.array_equal(osq.cpu().numpy(), ob.sumsq)
    assert cnt.cpu().numpy().tolist() == [float(ob.count), float(saved)]


# ------------------------------------------------------------------------------------------- a4 + a5
def _rollout_case(eng, obs_dim, act_dim, hidden, T, n_pairs, seed=0, sigma=0.02, with_stats=True):
    rs = np.random.RandomState(seed)
    dims = orc.layer_dims(obs_dim, hidden, act_dim)
    P = orc.n_params(dims)
    L = P + 50_000
    table = rs.randn(L).astype(np.float32)
    theta = (rs.randn(P) * 0.1).astype(np.float32)
    env = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    mean = rs.randn(obs_dim) * 0.1 if with_stats else np.zeros(obs_dim)
    std = 0.5 + rs.rand(obs_dim) if with_stats else np.ones(obs_dim)
    idx = rs.randint(0, L - P, size=n_pairs).astype(np.int64)
    idx[0] = 0
    idx[-1] = L - P - 1
    obsn = eng.normalise_obs(dev(eng, env.obs_stream[:T]), dev(eng, mean), dev(eng, std), 5.0)
    fit = torch.zeros(2, n_pairs, dtype=torch.float64, device=eng.device)
    behv = torch.zeros(2, n_pairs, 3, dtype=torch.float32, device=eng.device)
    eng.rollout(dev(eng, table), dev(eng, idx), dev(eng, theta), sigma, [obs_dim] + list(hidden) + [act_dim], obsn,
                dev(eng, env.rew_vec), env.pos_scale, fit[0], fit[1], 1, behv[0], behv[1])
    fit, behv = fit.cpu().numpy(), behv.cpu().numpy()
    for k in range(n_pairs):
        noise = orc.table_get(table, int(idx[k]), P)
        for s, nz in ((0, noise), (1, -noise)):
            layers = orc.unflatten(orc.pheno_params(theta, sigma, nz), dims)
            rews, b, _, _ = orc.run_model(env, layers, mean, std, 5.0, T, batched=True)
            ref = orc.reward_result(rews)[0]
            scale = max(1.0, np.abs(rews).sum())
            # float32 forward, different summation order than torch-CPU: 1e-5 of the episode's |reward| mass
            assert abs(fit[s, k] - ref) <= 1e-5 * scale, (k, s, fit[s, k], ref)
            assert np.allclose(behv[s, k], b[-3:], rtol=1e-4, atol=1e-5)


def test_rollout_f32_halfcheetah_shape(eng):
    _rollout_case(eng, 17, 6, (64, 64), T=100, n_pairs=6)


def test_rollout_f32_humanoid_shape(eng):
    _rollout_case(eng, 376, 17, (64, 64), T=70, n_pairs=4, seed=3)


def test_rollout_f32_odd_shapes(eng):
    _rollout_case(eng, 15, 3, (33,), T=33, n_pairs=3, seed=4)           # ragged: nothing is a multiple of 4
    _rollout_case(eng, 5, 1, (8, 8, 8), T=1, n_pairs=2, seed=5)         # single step, single action


def test_rollout_f32_large_network_simple_conf_shape(eng):
    """configs/simple_conf.json's policy (15 -> 256 -> 256 -> 3, P = 70 659): its padded weights (283 KB) do not fit in
    shared memory, the kernel reads them from a staged global scratch instead.  Same tolerance; also with a time split
    (1 pair) and with enough pairs to need several launches of staged weights."""
    _rollout_case(eng, 15, 3, (256, 256), T=40, n_pairs=3, seed=21)
    _rollout_case(eng, 15, 3, (256, 256), T=100, n_pairs=1, seed=22)


def test_rollout_f32_time_split(eng):
    """Fewer policies than SMs: the episode's time tiles are split over the SMs (single evaluations of the compatibility
    path, es.step's noiseless evaluation); same tolerance as the unsplit kernel."""
    _rollout_case(eng, 17, 6, (64, 64), T=1000, n_pairs=1, seed=11)
    _rollout_case(eng, 376, 17, (64, 64), T=333, n_pairs=2, seed=12)
    _rollout_case(eng, 17, 6, (64, 64), T=40, n_pairs=40, seed=13)       # 80 policies, 2 tiles: 1 split each


def test_rollout_f32x_vs_oracle(eng):
    """Enough pairs to fill the GPU (2 * n_pairs >= SM count): the packed-FMA kernel of rollout_f32x.cu (one CTA per pair,
    layer 1 shared by both signs as U +- sigma*V, accumulators paired along k).  Same oracle, same tolerance as the general
    kernel."""
    sms = torch.cuda.get_device_properties(eng.device).multi_processor_count
    n = (sms + 1) // 2
    _rollout_case(eng, 17, 6, (64, 64), T=150, n_pairs=n + 6, seed=31)       # 2 tiles (128 + 22 rows); obs not a multiple of 4
    _rollout_case(eng, 376, 17, (64, 64), T=130, n_pairs=n + 2, seed=32)     # the bench's shape
    _rollout_case(eng, 5, 1, (64, 64), T=3, n_pairs=n + 1, seed=33)          # one action, one partial tile
    _rollout_case(eng, 63, 32, (64, 64), T=129, n_pairs=2 * sms + 11, seed=34)   # CTAs walk several pairs; act = 32


def test_rollout_f32x_matches_general_kernel(eng, monkeypatch):
    """The two float32 kernels on identical inputs at a size the oracle loop would take minutes for (T = 1000): fitness
    within 2e-6 of the episode's |reward| mass of each other (both are within 1e-5 of the oracle), positions to float32
    rounding of a T-term sum."""
    rs = np.random.RandomState(77)
    obs, act, T, n = 376, 17, 1000, 300
    sizes = [obs, 64, 64, act]
    P = orc.n_params(orc.layer_dims(obs, (64, 64), act))
    L = P + 400_000
    table, theta = dev(eng, rs.randn(L).astype(np.float32)), dev(eng, (rs.randn(P) * 0.1).astype(np.float32))
    idx = dev(eng, rs.randint(0, L - P - 1, size=n).astype(np.int64))
    rew_h = rs.randn(T, act).astype(np.float32)
    obsn, rew = dev(eng, np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)), dev(eng, rew_h)
    out = {}
    for general in (False, True):
        if general:
            monkeypatch.setenv('ES_F32_GENERAL', '1')
        fit = torch.zeros(2, n, dtype=torch.float64, device=eng.device)
        behv = torch.zeros(2, n, 3, dtype=torch.float32, device=eng.device)
        eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], 1, behv[0], behv[1])
        eng.sync()
        out[general] = (fit.cpu().numpy(), behv.cpu().numpy())
    (fx, bx), (fg, bg) = out[False], out[True]
    assert not np.array_equal(fx, fg), 'ES_F32_GENERAL did not select the other kernel'
    assert np.abs(fx - fg).max() <= 2e-6 * np.abs(rew_h).sum()
    assert np.abs(bx - bg).max() <= 1e-5


def test_rollout_sigma_zero_is_symmetric(eng):
    rs = np.random.RandomState(9)
    sizes = [17, 64, 64, 6]
    P = orc.n_params(orc.layer_dims(17, (64, 64), 6))
    env = orc.SyntheticEnvSpec(17, 6, 64)
    table, theta = dev(eng, rs.randn(P + 1000).astype(np.float32)), dev(eng, (rs.randn(P) * .1).astype(np.float32))
    idx = dev(eng, rs.randint(0, 1000, size=16).astype(np.int64))
    fit = torch.zeros(2, 16, dtype=torch.float64, device=eng.device)
    eng.rollout(table, idx, theta, 0.0, sizes, dev(eng, env.obs_stream[:64]), dev(eng, env.rew_vec), 0.05, fit[0], fit[1])
    f = fit.cpu().numpy()
    assert np.array_equal(f[0], f[1]) and np.all(f[0] == f[0][0])


def test_rollout_too_wide_fails_loudly(eng):
    """Weights larger than shared memory are staged in global memory (test_rollout_f32_large_network_...); a layer so wide that
    the activation tiles themselves do not fit is refused with a message, not computed somewhere else."""
    from es_pytorch_b200._lib import EsLibraryError
    sizes = [15, 1024, 3]                           # 2 x 32 x 1024 float32 activation tiles = 262 kB > 227 kB
    P = orc.n_params(orc.layer_dims(15, (1024,), 3))
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=eng.device)
    fit = torch.zeros(2, 1, dtype=torch.float64, device=eng.device)
    with pytest.raises(EsLibraryError, match='shared memory'):
        eng.rollout(z(P + 10), torch.zeros(1, dtype=torch.int64, device=eng.device), z(P), 0.02, sizes, z(8, 15),
                    z(8, 3), 0.05, fit[0], fit[1])


# ------------------------------------------------------------------------------------------- a8 / a9
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_rank_golden(eng, ref_vectors, tag):
    v = ref_vectors
    w = eng.centered_rank(dev(eng, v[f'rank1_{tag}_pos']), dev(eng, v[f'rank1_{tag}_neg']))
    assert np.array_equal(w.cpu().numpy(), v[f'rank1_{tag}_w'])
    for wtag, wt in (('w03', 0.3), ('w10', 1.0), ('w00', 0.0)):
        w2 = eng.centered_rank(dev(eng, v[f'rank2_{tag}_pos']), dev(eng, v[f'rank2_{tag}_neg']), wt, 1 - wt)
        assert np.array_equal(w2.cpu().numpy(), v[f'rank2_{tag}_{wtag}_w'])


@pytest.mark.parametrize('K,n_obj', [(1, 1), (2, 2), (777, 1), (3000, 2), (10000, 1)])
def test_rank_vs_oracle_and_shards(eng, K, n_obj):
    rs = np.random.RandomState(K)
    pos, neg = rs.randn(K, n_obj) * 10, rs.randn(K, n_obj) * 10
    if K > 100:                                      # inject ties, signed zeros and infinities
        pos[5] = pos[17]; neg[3] = pos[5]; pos[40] = 0.0; neg[41] = -0.0; pos[60] = np.inf; neg[61] = -np.inf
    ref = orc.centered_ranker(pos, neg)[0] if n_obj == 1 else orc.moo_ranker(pos, neg, 0.37)[0]
    w0, w1 = (1.0, 0.0) if n_obj == 1 else (0.37, 1 - 0.37)
    fp, fn = dev(eng, pos), dev(eng, neg)
    w, ranks = eng.centered_rank(fp, fn, w0, w1, want_ranks=True)
    assert np.array_equal(w.cpu().numpy(), ref)
    full = np.concatenate((pos, neg))
    for c in range(n_obj):
        r = ranks[c].cpu().numpy().ravel()
        assert np.array_equal(r, orc.rank(full[:, c]))
    # a GPU's shard gets exactly its slice of the global weights
    if K >= 4:
        b, cnt = K // 4, K // 2
        ws = eng.centered_rank(fp, fn, w0, w1, k_begin=b, k_count=cnt)
        assert np.array_equal(ws.cpu().numpy(), ref[b:b + cnt])


def test_rank_integer_ties_and_nan(eng):
    pos = np.array([[3.], [1.], [3.], [0.], [2.], [np.nan]])
    neg = np.array([[1.], [3.], [2.], [2.], [-0.], [5.]])
    w, ranks = eng.centered_rank(dev(eng, pos), dev(eng, neg), want_ranks=True)
    x = np.concatenate((pos, neg)).ravel()
    assert np.array_equal(ranks.cpu().numpy().ravel(), orc.rank(x))       # NaN last, ties by position
    assert np.array_equal(w.cpu().numpy(), orc.centered_ranker(pos, neg)[0])


# ------------------------------------------------------------------------------------------- f4: rankers.py:61-103
RANKER_CLASSES = {'centered': 'CenteredRanker', 'double_positive': 'DoublePositiveCenteredRanker',
                  'max_normalized': 'MaxNormalizedRanker', 'semi_centered': 'SemiCenteredRanker'}


def _ranker(name):
    from es_pytorch_b200.utils import rankers as R
    return getattr(R, RANKER_CLASSES[name])()


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
@pytest.mark.parametrize('name', ['double_positive', 'max_normalized', 'semi_centered'])
def test_shaped_rankers_golden(eng, ranker_vectors, tag, name):
    """The public Ranker API against outputs of the real reference module: bit-exact, dtype and shape included."""
    from es_pytorch_b200.utils import rankers as R
    v = ranker_vectors
    r = _ranker(name)
    w = r.rank(v[f'{tag}_pos'], v[f'{tag}_neg'], v[f'{tag}_inds'])
    ref = v[f'{tag}_{name}_w']
    assert w.dtype == ref.dtype and w.shape == ref.shape and r.n_fits_ranked == int(v[f'{tag}_{name}_n'])
    assert np.array_equal(w, ref)
    m = R.MultiObjectiveRanker(_ranker(name), 0.3)
    w2 = m.rank(v[f'{tag}_pos2'], v[f'{tag}_neg2'], v[f'{tag}_inds'])
    ref2 = v[f'{tag}_moo_{name}_w']
    assert w2.dtype == ref2.dtype and np.array_equal(w2.reshape(ref2.shape), ref2)


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
@pytest.mark.parametrize('name', ['centered', 'double_positive', 'max_normalized'])
@pytest.mark.parametrize('ptag,pct', [('p00', 0.0), ('p10', 0.1), ('p50', 0.5), ('p100', 1.0)])
def test_elite_ranker_golden(eng, ranker_vectors, tag, name, ptag, pct):
    from es_pytorch_b200.utils import rankers as R
    v = ranker_vectors
    e = R.EliteRanker(_ranker(name), pct)
    vals = e.rank(v[f'{tag}_pos'], v[f'{tag}_neg'], v[f'{tag}_inds'])
    ref_v = v[f'{tag}_elite_{name}_{ptag}_vals']
    assert e.n_fits_ranked == int(v[f'{tag}_elite_{name}_{ptag}_n']) == len(vals) == len(e.noise_inds)
    assert vals.dtype == ref_v.dtype
    order = np.lexsort((e.noise_inds, vals))            # np.argpartition's order is unspecified: compare as sets
    assert np.array_equal(np.asarray(vals)[order], ref_v)
    assert np.array_equal(np.asarray(e.noise_inds)[order], v[f'{tag}_elite_{name}_{ptag}_inds'])


@pytest.mark.parametrize('name', ['centered', 'double_positive', 'max_normalized', 'semi_centered'])
@pytest.mark.parametrize('K,n_obj', [(1, 1), (777, 1), (3000, 2), (10000, 1)])
def test_rank_transform_vs_oracle_and_shards(eng, name, K, n_obj):
    from es_pytorch_b200 import _lib
    kind = getattr(_lib, 'ES_RANK_' + name.upper())
    rs = np.random.RandomState(K + 1)
    pos, neg = rs.randn(K, n_obj) * 10 + 1, rs.randn(K, n_obj) * 10 + 1
    if K > 100:
        pos[5] = pos[17]; neg[3] = pos[5]; pos[40] = 0.0; neg[41] = -0.0     # ties and signed zeros
    if K == 1 and name == 'max_normalized':
        pytest.skip('a single pair normalises by max(y) = 0 for some draws')
    ref, n = orc.shaped_ranker(pos, neg, name, None if n_obj == 1 else 0.37)
    w0, w1 = (1.0, 0.0) if n_obj == 1 else (0.37, 1 - 0.37)
    fp, fn = dev(eng, pos), dev(eng, neg)
    out = eng.rank_transform(fp, fn, kind, w0, w1, want64=True)
    got = out['weights64'].cpu().numpy()
    assert np.array_equal(got, np.asarray(ref, dtype=np.float64).reshape(-1))     # float32 shapings widen exactly
    assert np.array_equal(out['weights'].cpu().numpy(), got.astype(np.float32))
    if K >= 4:
        b, cnt = K // 4, K // 2
        ws = eng.rank_transform(fp, fn, kind, w0, w1, k_begin=b, k_count=cnt, want64=True)['weights64']
        assert np.array_equal(ws.cpu().numpy(), got[b:b + cnt])


@pytest.mark.parametrize('name', ['centered', 'double_positive', 'max_normalized'])
@pytest.mark.parametrize('K,pct', [(777, 0.05), (10000, 0.1), (10000, 1.0)])
def test_elite_vs_oracle_pair_weights_and_shards(eng, name, K, pct):
    """Compact elite lists == oracle (same order: ascending shaped value); the per-pair weights the sharded generation
    uses are the scatter of that list; shards write disjoint parts of it."""
    from es_pytorch_b200 import _lib
    kind = getattr(_lib, 'ES_RANK_' + name.upper())
    rs = np.random.RandomState(K + 7)
    pos, neg = rs.randn(K, 1) * 10 + 3, rs.randn(K, 1) * 10 + 3
    inds = rs.randint(0, 10 ** 8, K).astype(np.int64)
    vals, sel, fit, n_el = orc.elite_ranker(pos, neg, inds, name, pct)
    fp, fn, di = dev(eng, pos), dev(eng, neg), dev(eng, inds)
    out = eng.rank_transform(fp, fn, kind, elite_n=n_el, noise_idx=di, want64=True, want_elite=True)
    assert np.array_equal(out['elite_fit'].cpu().numpy(), fit)
    assert np.array_equal(out['elite_idx'].cpu().numpy(), sel)
    assert np.array_equal(out['elite_vals'].cpu().numpy(), np.asarray(vals, dtype=np.float64))
    dt = np.float64 if name == 'max_normalized' else np.float32
    pw = np.zeros(K, dtype=dt)
    for vv, f in zip(np.asarray(vals, dtype=dt), fit):             # at most two terms per pair: order-free
        pw[f % K] += vv
    assert np.array_equal(out['weights64'].cpu().numpy(), pw.astype(np.float64))
    b, cnt = K // 4, K // 2
    part = eng.rank_transform(fp, fn, kind, elite_n=n_el, k_begin=b, k_count=cnt, noise_idx=di, want_elite=True)
    inside = (fit % K >= b) & (fit % K < b + cnt)
    assert np.array_equal(part['elite_fit'].cpu().numpy()[inside], fit[inside])
    assert np.all(part['elite_fit'].cpu().numpy()[~inside] == 0)
    assert np.array_equal(part['weights'].cpu().numpy(), pw[b:b + cnt].astype(np.float32))


def test_rank_transform_rejects(eng):
    from es_pytorch_b200._lib import EsLibraryError
    from es_pytorch_b200.utils import rankers as R
    fp = dev(eng, np.zeros((4, 2)))
    with pytest.raises(EsLibraryError):
        eng.rank_transform(fp, fp, 0, 0.5, 0.5, elite_n=2)         # elite over two objectives
    with pytest.raises(EsLibraryError):
        eng.rank_transform(fp, fp, 7)                              # unknown shaping
    with pytest.raises(NotImplementedError):
        R.EliteRanker(R.MultiObjectiveRanker(R.CenteredRanker(), 0.5), 0.1)
    with pytest.raises(ValueError):
        R.CenteredRanker().rank(np.zeros((4, 2)), np.zeros((4, 2)), np.arange(4))


def test_elite_approx_grad_matches_oracle(eng):
    """obj.py:50's EliteRanker(CenteredRanker(), elite) through es.approx_grad: theta within 1e-5 of the oracle."""
    from es_pytorch_b200.core import es
    from es_pytorch_b200.core.noisetable import NoiseTable
    from es_pytorch_b200.core.policy import Policy
    from es_pytorch_b200.nn.nn import FeedForward
    from es_pytorch_b200.nn.optimizers import Adam
    from es_pytorch_b200.gym import synthetic_env
    from es_pytorch_b200.utils import rankers as R
    import torch.nn as tnn
    env = synthetic_env.SyntheticEnv(17, 6, 20)
    net = FeedForward([64, 64], tnn.Tanh(), env, 0.0)
    P = len(Policy.get_flat(net))
    table = np.random.RandomState(5).randn(300_000).astype(np.float32)
    nt = NoiseTable(P, table)
    policy = Policy(net, 0.02, Adam(P, 0.01))
    theta0 = policy.flat_params.copy()
    K = 400
    rs = np.random.RandomState(9)
    pos, neg = rs.randn(K, 1), rs.randn(K, 1)
    inds = rs.randint(0, len(table) - P, K).astype(np.float64)
    ranker = R.EliteRanker(R.CenteredRanker(), 0.2)
    ranker.rank(pos, neg, inds)
    es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
    vals, sel, fit, n_el = orc.elite_ranker(pos, neg, inds, 'centered', 0.2)
    flat = theta0.copy()
    orc.approx_grad(flat, orc.AdamOracle(P, 0.01), vals, sel, n_el, table, 500, 0.005)
    assert ranker.n_fits_ranked == n_el
    assert np.max(np.abs(policy.flat_params - flat)) <= 1e-5 * max(1.0, np.max(np.abs(flat)))


# ------------------------------------------------------------------------------------------- a10
def test_reconstruct_reference_known_answer(eng):
    """test/utils/utils_test.py:24-40 (integer data: exact in any summation order)."""
    evals, params = 100, 500
    fits, inds, table = np.arange(evals), np.arange(evals), np.arange(2000)
    expected = np.dot(fits, [[i + j for j in range(params)] for i in range(evals)])
    out = eng.grad_reconstruct(dev(eng, table.astype(np.float32)), dev(eng, inds.astype(np.int64)),
                               dev(eng, fits.astype(np.float32)), params)
    assert np.array_equal(out.cpu().numpy(), expected.astype(np.float32))


@pytest.mark.parametrize('K,P', [(1, 7), (37, 1023), (256, 5702), (1000, 29393), (5, 70659)])
def test_reconstruct_vs_oracle(eng, K, P):
    rs = np.random.RandomState(K + P)
    L = P + 300_000
    table = rs.randn(L).astype(np.float32)
    idx = rs.randint(0, L - P, size=K).astype(np.int64)
    idx[0] = L - P - 1                                   # last admissible slice (noisetable.py:34)
    if K > 2:
        idx[1] = idx[2]                                  # duplicate index (es.py:44 reports dupes)
    w = (rs.rand(K).astype(np.float32) - 0.5) * 2
    out = eng.grad_reconstruct(dev(eng, table), dev(eng, idx), dev(eng, w), P).cpu().numpy()
    ref32 = np.asarray(orc.scale_noise(w, idx.astype(np.float64), table, P, 500), dtype=np.float32)
    truth = orc.scale_noise_f64(w, idx, table, P)
    scale = np.abs(truth).max()
    assert np.abs(out - truth).max() <= 1e-5 * scale                 # north-star tolerance, vs float64 truth
    assert np.abs(out - ref32).max() <= 1e-5 * scale                 # and vs the reference's float32 sgemv order
    out2 = eng.grad_reconstruct(dev(eng, table), dev(eng, idx), dev(eng, w), P).cpu().numpy()
    assert np.array_equal(out, out2)                                 # deterministic


def test_reconstruct_empty_and_onehot(eng):
    rs = np.random.RandomState(5)
    P, L = 2000, 50_000
    table = rs.randn(L).astype(np.float32)
    t = dev(eng, table)
    out = eng.grad_reconstruct(t, dev(eng, np.zeros(0, dtype=np.int64)), dev(eng, np.zeros(0, dtype=np.float32)), P)
    assert not out.cpu().numpy().any()
    idx = rs.randint(0, L - P, size=300).astype(np.int64)
    w = np.zeros(300, dtype=np.float32)
    w[123] = 1.0
    out = eng.grad_reconstruct(t, dev(eng, idx), dev(eng, w), P).cpu().numpy()
    assert np.array_equal(out, table[idx[123]:idx[123] + P])        # a one-hot weight returns the slice itself


# ------------------------------------------------------------------------------------------- a11 / a12
@pytest.mark.parametrize('kind', ['adam', 'sgd', 'simple'])
def test_optimizer_steps_bit_exact(eng, kind):
    from es_pytorch_b200.nn import optimizers as O
    rs = np.random.RandomState(11)
    P, K = 5702, 256
    theta = (rs.randn(P) * 0.1).astype(np.float32)
    mine = {'adam': O.Adam(P, 0.01), 'sgd': O.SGD(P, 0.02), 'simple': O.SimpleES(P, 0.03)}[kind]
    ref = {'adam': orc.AdamOracle(P, 0.01), 'sgd': orc.SGDOracle(P, 0.02), 'simple': orc.SimpleESOracle(P, 0.03)}[kind]
    th_dev, th_ref = dev(eng, theta), theta.copy()
    for it in range(5):
        gsum = (rs.randn(P) * 3).astype(np.float32)
        mine.apply_fused(eng, th_dev, dev(eng, gsum), float(2 * K), 0.005)
        grad = (gsum / np.float32(2 * K)).astype(np.float32)
        g = ((np.float32(0.005) * th_ref).astype(np.float32) - grad).astype(np.float32)
        th_ref += ref.step(g)
        assert np.array_equal(th_dev.cpu().numpy(), th_ref), it
    if kind == 'adam':
        assert np.array_equal(mine.m, ref.m) and np.array_equal(mine.v, ref.v) and mine.t == ref.t == 5
    # reference-style step(g) -> delta contract
    g = rs.randn(P).astype(np.float32)
    assert np.array_equal(mine.step(g), ref.step(g))


def test_optimizers_golden(eng, ref_vectors):
    from es_pytorch_b200.nn import optimizers as O
    v = ref_vectors
    sgd, ses, adam = O.SGD(40, 0.01), O.SimpleES(40, 0.01), O.Adam(40, 0.01)
    for i, g in enumerate(v['sgd_g']):
        assert np.array_equal(sgd.step(g), v['sgd_steps'][i])           # real reference module, bit-exact
        assert np.array_equal(ses.step(g), v['simple_steps'][i])
        assert np.allclose(adam.step(g), v['adam_steps_real_f64'][i], rtol=2e-6, atol=1e-9)


# ------------------------------------------------------------------------------------------- a13
def test_novelty(eng):
    rs = np.random.RandomState(21)
    n, A = 500, 64
    behv = rs.randn(n, 3).astype(np.float32)
    archive = rs.randn(A, 2)
    for k in (1, 10, 64, 100):
        out = torch.zeros(n, dtype=torch.float64, device=eng.device)
        eng.novelty(dev(eng, behv), dev(eng, archive), k, out)
        ref = np.array([orc.novelty(behv[i, :2], archive, k) for i in range(n)])
        assert np.allclose(out.cpu().numpy(), ref, rtol=1e-14, atol=0)
    # reference known answers (test/utils/novelty_test.py:27-33)
    b = dev(eng, np.zeros((1, 3), dtype=np.float32))
    arch = dev(eng, np.array([[2., 2.], [1., 1.], [3., 3.]]))
    for k, expect in ((1, np.sqrt(2)), (2, (np.sqrt(2) + np.sqrt(8)) / 2), (3, (np.sqrt(2) + np.sqrt(8) + np.sqrt(18)) / 3),
                      (50, (np.sqrt(2) + np.sqrt(8) + np.sqrt(18)) / 3)):
        out = torch.zeros(1, dtype=torch.float64, device=eng.device)
        eng.novelty(b, arch, k, out)
        assert out.item() == expect


# ------------------------------------------------------------------------------------------- tensor-core rollouts
def _tc_vs_f32(eng, obs, act, T, n_pairs, seed, mode):
    from es_pytorch_b200 import _lib
    rs = np.random.RandomState(seed)
    sizes = [obs, 64, 64, act]
    P = orc.n_params(orc.layer_dims(obs, (64, 64), act))
    L = P + 1_000_000
    table, theta = dev(eng, rs.randn(L).astype(np.float32)), dev(eng, (rs.randn(P) * 0.1).astype(np.float32))
    idx = dev(eng, rs.randint(0, L - P - 1, size=n_pairs).astype(np.int64))
    obsn_h, rew_h = np.clip(rs.randn(T, obs), -5, 5).astype(np.float32), rs.randn(T, act).astype(np.float32)
    obsn, rew = dev(eng, obsn_h), dev(eng, rew_h)
    res = {}
    for md in (_lib.ES_ROLLOUT_F32, mode):
        fit = torch.zeros(2, n_pairs, dtype=torch.float64, device=eng.device)
        behv = torch.zeros(2, n_pairs, 3, dtype=torch.float32, device=eng.device)
        eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], 1, behv[0], behv[1], md)
        eng.sync()
        res[md] = (fit.cpu().numpy(), behv.cpu().numpy())
    return res[_lib.ES_ROLLOUT_F32], res[mode], np.abs(rew_h).sum()


_TC_SHAPES = [(376, 17, 1000, 200), (17, 6, 1000, 256), (17, 6, 100, 8), (5, 1, 130, 3), (63, 32, 129, 5), (64, 3, 128, 4),
              (24, 9, 257, 150),                      # TMA path (obs % 8 == 0), action columns split 8 + 1 over the two halves
              # several pairs per CTA with an odd tile count (the two epilogue groups swap parity every pair) and with a
              # single tile per pair (one group per pair)
              (17, 6, 300, 333), (17, 6, 100, 400)]


@pytest.mark.parametrize('obs,act,T,n_pairs', _TC_SHAPES)
def test_rollout_tc3_is_float32_equivalent(eng, obs, act, T, n_pairs):
    """ES_ROLLOUT_TC3 (float16 hi+lo split operands, three tcgen05 MMAs per product, accurate tanh, float64 sums) against
    the float32 CUDA-core rollout (itself within 1e-5 of the oracle above): same tolerance class as that test -- the
    fitness differs by less than 1e-5 of the episode's |reward| mass and by a few 1e-6 of the population's fitness spread
    (measured: 1.4e-6 .. 1.9e-6 rms), where the float16 single-product path is ~1e-3."""
    from es_pytorch_b200 import _lib
    (f32, b32), (ftc, btc), mass = _tc_vs_f32(eng, obs, act, T, n_pairs, obs + T, _lib.ES_ROLLOUT_TC3)
    assert np.abs(ftc - f32).max() <= 1e-5 * max(1.0, mass / 8), (np.abs(ftc - f32).max(), mass)
    spread = max(f32.std(), 1e-3 * np.sqrt(T))
    assert np.sqrt(((ftc - f32) ** 2).mean()) <= 6e-6 * spread + 1e-6
    assert np.abs(btc - b32).max() <= 2e-6 * 0.05 * T + 1e-6             # final positions: float32 sums of T terms
    if n_pairs >= 100:                                                   # ranks: only adjacent near-ties may swap
        r32, rtc = np.argsort(np.argsort(f32.ravel())), np.argsort(np.argsort(ftc.ravel()))
        assert np.abs(r32 - rtc).max() <= 2 and (r32 != rtc).mean() <= 0.02


@pytest.mark.parametrize('obs,act,T,n_pairs', _TC_SHAPES)
def test_rollout_tc_matches_f32(eng, obs, act, T, n_pairs):
    """ES_ROLLOUT_TC (float16 operands, one MMA per product, tanh.approx) vs the float32 CUDA-core rollout: fitness within
    ~1e-3 of the population's fitness spread (measured 0.6e-3 .. 2.3e-3 rms), ranks essentially unchanged."""
    from es_pytorch_b200 import _lib
    (f32, b32), (ftc, btc), _ = _tc_vs_f32(eng, obs, act, T, n_pairs, obs + T, _lib.ES_ROLLOUT_TC)
    spread = max(f32.std(), 1e-3 * np.sqrt(T))
    assert np.abs(ftc - f32).max() <= 0.02 * spread + 1e-3 * np.sqrt(T) * 0.05
    assert np.sqrt(((ftc - f32) ** 2).mean()) <= 5e-3 * spread + 2e-4
    d32, dtc = f32[0] - f32[1], ftc[0] - ftc[1]
    assert np.sqrt(((dtc - d32) ** 2).mean()) <= 5e-3 * max(d32.std(), 1e-3 * np.sqrt(T)) + 1e-3
    assert np.abs(btc - b32).max() <= 2e-3 * 0.05 * T + 1e-4
    if n_pairs >= 100:
        r32, rtc = np.argsort(np.argsort(f32.ravel())), np.argsort(np.argsort(ftc.ravel()))
        assert np.corrcoef(r32, rtc)[0, 1] > 0.99999


def test_rollout_tc3_error_against_float64_truth(eng):
    """What "float32-equivalent" means, measured: on the Humanoid-shaped policy (T = 1000) the split tensor-core path's
    fitness error against float64 arithmetic on the same inputs is within 4x of the float32 CUDA-core path's own error
    (measured 2.4x: 1.8e-6 vs 0.8e-6 of the fitness spread), three orders of magnitude below the single-product path."""
    from es_pytorch_b200 import _lib
    rs = np.random.RandomState(376 + 1000)
    obs, act, T, n = 376, 17, 1000, 24
    sizes = [obs, 64, 64, act]
    P = orc.n_params(orc.layer_dims(obs, (64, 64), act))
    L = P + 1_000_000
    table_h, theta_h = rs.randn(L).astype(np.float32), (rs.randn(P) * 0.1).astype(np.float32)
    idx_h = rs.randint(0, L - P - 1, size=n).astype(np.int64)
    obsn_h, rew_h = np.clip(rs.randn(T, obs), -5, 5).astype(np.float32), rs.randn(T, act).astype(np.float32)
    truth = np.zeros((2, n))
    x, c = obsn_h.astype(np.float64), rew_h.astype(np.float64)
    for k, i in enumerate(idx_h):
        for s, sign in enumerate((1.0, -1.0)):
            w = theta_h.astype(np.float64) + sign * np.float64(np.float32(0.02)) * table_h[i:i + P].astype(np.float64)
            a, at = x, 0
            for fi, fo in ((obs, 64), (64, 64), (64, act)):
                W = w[at:at + fi * fo].reshape(fo, fi); at += fi * fo
                b = w[at:at + fo]; at += fo
                a = np.tanh(a @ W.T + b)
            truth[s, k] = (a * c).sum()
    err = {}
    for mode in (_lib.ES_ROLLOUT_F32, _lib.ES_ROLLOUT_TC3, _lib.ES_ROLLOUT_TC):
        fit = torch.zeros(2, n, dtype=torch.float64, device=eng.device)
        eng.rollout(dev(eng, table_h), dev(eng, idx_h), dev(eng, theta_h), 0.02, sizes, dev(eng, obsn_h), dev(eng, rew_h), 0.05,
                    fit[0], fit[1], mode=mode)
        eng.sync()
        err[mode] = np.sqrt(((fit.cpu().numpy() - truth) ** 2).mean())
    spread = truth.std()
    assert err[_lib.ES_ROLLOUT_F32] <= 3e-6 * spread
    assert err[_lib.ES_ROLLOUT_TC3] <= 4 * err[_lib.ES_ROLLOUT_F32] and err[_lib.ES_ROLLOUT_TC3] <= 6e-6 * spread
    assert err[_lib.ES_ROLLOUT_TC] >= 50 * err[_lib.ES_ROLLOUT_TC3]


@pytest.mark.parametrize('tc_mode', [1, 2])
def test_rollout_tc_shadow_tracks_table_contents(eng, tc_mode):
    """The tensor-core paths read layer 1 from float16 shadows of the table (obs % 8 == 0): an in-place rewrite of the table
    tensor, or a new tensor at a recycled address, must be picked up; obs % 8 != 0 takes the float32-slice path."""
    from es_pytorch_b200 import _lib
    assert (_lib.ES_ROLLOUT_TC, _lib.ES_ROLLOUT_TC3) == (1, 2)
    rs = np.random.RandomState(5)
    obs, act, T, n = 24, 6, 100, 40
    sizes = [obs, 64, 64, act]
    P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
    L = P + 30_000
    theta = dev(eng, (rs.randn(P) * 0.1).astype(np.float32))
    idx = dev(eng, rs.randint(0, L - P, size=n).astype(np.int64))
    obsn = dev(eng, np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)); rew = dev(eng, rs.randn(T, act).astype(np.float32))

    def run(table, mode):
        fit = torch.zeros(2, n, dtype=torch.float64, device=eng.device)
        eng.rollout(table, idx, theta, 0.05, sizes, obsn, rew, 0.05, fit[0], fit[1], mode=mode)
        return fit.cpu().numpy()

    def close(a, b):
        return np.abs(a - b).max() <= 0.02 * max(b.std(), 1e-2)

    t1 = dev(eng, rs.randn(L).astype(np.float32))
    a_tc, a_32 = run(t1, tc_mode), run(t1, _lib.ES_ROLLOUT_F32)
    assert close(a_tc, a_32)
    t1.copy_(torch.from_numpy(rs.randn(L).astype(np.float32)))               # in place: same pointer, new contents
    b_tc, b_32 = run(t1, tc_mode), run(t1, _lib.ES_ROLLOUT_F32)
    assert close(b_tc, b_32) and not close(b_tc, a_32)
    del t1
    t2 = dev(eng, rs.randn(L).astype(np.float32))                             # usually the recycled address of t1
    c_tc, c_32 = run(t2, tc_mode), run(t2, _lib.ES_ROLLOUT_F32)
    assert close(c_tc, c_32) and not close(c_tc, b_32)


@pytest.mark.parametrize('tc_mode', [1, 2])
def test_rollout_tc_sigma_zero_symmetric_and_unsupported_shape(eng, tc_mode):
    from es_pytorch_b200 import _lib
    from es_pytorch_b200._lib import EsLibraryError
    rs = np.random.RandomState(3)
    P = orc.n_params(orc.layer_dims(17, (64, 64), 6))
    table, theta = dev(eng, rs.randn(P + 5000).astype(np.float32)), dev(eng, (rs.randn(P) * .1).astype(np.float32))
    idx = dev(eng, rs.randint(0, 5000, size=200).astype(np.int64))
    obsn, rew = dev(eng, rs.randn(256, 17).astype(np.float32)), dev(eng, rs.randn(256, 6).astype(np.float32))
    fit = torch.zeros(2, 200, dtype=torch.float64, device=eng.device)
    eng.rollout(table, idx, theta, 0.0, [17, 64, 64, 6], obsn, rew, 0.05, fit[0], fit[1], mode=tc_mode)
    f = fit.cpu().numpy()
    assert np.array_equal(f[0], f[1]) and np.all(f[0] == f[0][0])        # sigma = 0: every evaluation identical
    with pytest.raises(EsLibraryError, match='tensor-core path'):
        P2 = orc.n_params(orc.layer_dims(17, (32,), 6))
        eng.rollout(table, idx, dev(eng, np.zeros(P2, dtype=np.float32)), 0.02, [17, 32, 6], obsn, rew, 0.05, fit[0], fit[1],
                    mode=tc_mode)
