"""Generates tests/golden/*.npz.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

Two kinds of vectors:
  * ``ref_*``  -- outputs of the REAL reference modules that import here
    (src.utils.rankers, src.nn.optimizers) and of numpy's legacy RandomState (the pinned
    third-party RNG the reference draws indices from);
  * ``orc_*``  -- outputs of oracle/es_oracle.py for the parts no reference test pins
    (pheno, forward, rollout, a whole small generation), frozen so that a later oracle edit
    cannot silently move the target.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import es_oracle as orc  # noqa: E402


def ref_vectors():
    sys.path.insert(0, '/root/reference')
    from src.utils import rankers as R
    from src.nn import optimizers as O
    out = {}
    rs = np.random.RandomState(2024)
    for tag, k in (('a', 7), ('b', 64), ('c', 501)):
        pos, neg = rs.randn(k, 1) * 3, rs.randn(k, 1) * 3
        cr = R.CenteredRanker()
        out[f'rank1_{tag}_pos'], out[f'rank1_{tag}_neg'] = pos, neg
        out[f'rank1_{tag}_w'] = cr.rank(pos, neg, np.arange(k))
        out[f'rank1_{tag}_n'] = np.array(cr.n_fits_ranked)
        pos2, neg2 = rs.randn(k, 2), rs.randn(k, 2)
        for wtag, w in (('w03', 0.3), ('w10', 1.0), ('w00', 0.0)):
            mr = R.MultiObjectiveRanker(R.CenteredRanker(), w)
            out[f'rank2_{tag}_{wtag}_w'] = mr.rank(pos2, neg2, np.arange(k))
        out[f'rank2_{tag}_pos'], out[f'rank2_{tag}_neg'] = pos2, neg2
    # small integer data with ties: numpy sorts n < 16 by insertion (stable) so the tie order is defined
    pos, neg = np.array([[3.], [1.], [3.], [0.], [2.]]), np.array([[1.], [3.], [2.], [2.], [-0.]])
    out['rank_ties_pos'], out['rank_ties_neg'] = pos, neg
    out['rank_ties_w'] = R.CenteredRanker().rank(pos, neg, np.arange(5))
    # optimizers: SGD is float32 under any numpy; Adam's real module computes float64 under numpy 2
    g = rs.randn(6, 40).astype(np.float32)
    sgd = O.SGD(40, 0.01)
    out['sgd_g'] = g
    out['sgd_steps'] = np.stack([sgd.step(x) for x in g])
    adam = O.Adam(40, 0.01)
    out['adam_steps_real_f64'] = np.stack([adam.step(x) for x in g]).astype(np.float64)
    ses = O.SimpleES(40, 0.01)
    out['simple_steps'] = np.stack([ses.step(x) for x in g])
    # legacy RandomState streams (numpy frozen MT19937 + masked rejection)
    for tag, seed, n, ub, extra in (('a', 1000, 300, 250_000_000 - 29393, 4), ('b', 1001, 900, 1000, 0),
                                    ('c', 7, 257, (1 << 31) + 5, 2), ('d', 8, 64, 3, 1)):
        r = np.random.RandomState(seed)
        st = r.get_state()
        idx, ext = [], []
        for _ in range(n):
            idx.append(int(r.randint(0, ub)))
            ext.append([int.from_bytes(r.bytes(4), 'little') for _ in range(extra)])
        out[f'mt_{tag}_key0'], out[f'mt_{tag}_pos0'] = st[1].astype(np.uint32), np.array(st[2])
        out[f'mt_{tag}_cfg'] = np.array([seed, n, ub, extra], dtype=np.int64)
        out[f'mt_{tag}_idx'] = np.array(idx, dtype=np.int64)
        out[f'mt_{tag}_extra'] = np.array(ext, dtype=np.uint32).reshape(n, extra)
        out[f'mt_{tag}_key1'], out[f'mt_{tag}_pos1'] = r.get_state()[1].astype(np.uint32), np.array(r.get_state()[2])
    np.savez_compressed(os.path.join(HERE, 'ref_vectors.npz'), **out)
    print('ref_vectors.npz', len(out), 'arrays')


def ref_ranker_vectors():
    """The rankers outside the north-star pair (rankers.py:61-103), from the REAL module."""
    sys.path.insert(0, '/root/reference')
    from src.utils import rankers as R
    out = {}
    rs = np.random.RandomState(4242)
    plain = {'double_positive': R.DoublePositiveCenteredRanker, 'max_normalized': R.MaxNormalizedRanker,
             'semi_centered': R.SemiCenteredRanker, 'centered': R.CenteredRanker}
    for tag, k, off in (('a', 7, 0.0), ('b', 64, 0.0), ('c', 501, 0.0), ('d', 33, 20.0)):
        pos, neg = rs.randn(k, 1) * 3 + off, rs.randn(k, 1) * 3 + off
        pos2, neg2 = rs.randn(k, 2) + off, rs.randn(k, 2) + off
        inds = rs.randint(0, 10 ** 6, k).astype(np.float64)
        out[f'{tag}_pos'], out[f'{tag}_neg'], out[f'{tag}_pos2'], out[f'{tag}_neg2'] = pos, neg, pos2, neg2
        out[f'{tag}_inds'] = inds
        for name, cls in plain.items():
            if name != 'centered':
                r = cls()
                out[f'{tag}_{name}_w'] = r.rank(pos, neg, inds)
                out[f'{tag}_{name}_n'] = np.array(r.n_fits_ranked)
                m = R.MultiObjectiveRanker(cls(), 0.3)
                out[f'{tag}_moo_{name}_w'] = m.rank(pos2, neg2, inds)
            if name == 'semi_centered':
                continue        # EliteRanker(SemiCenteredRanker) fails inside the reference (argpartition on a [2K,1] array)
            for ptag, pct in (('p00', 0.0), ('p10', 0.1), ('p50', 0.5), ('p100', 1.0)):
                e = R.EliteRanker(cls(), pct)
                vals = e.rank(pos, neg, inds)
                order = np.lexsort((e.noise_inds, vals))            # argpartition order is unspecified: store sorted
                out[f'{tag}_elite_{name}_{ptag}_vals'] = np.asarray(vals)[order]
                out[f'{tag}_elite_{name}_{ptag}_inds'] = np.asarray(e.noise_inds)[order]
                out[f'{tag}_elite_{name}_{ptag}_n'] = np.array(e.n_fits_ranked)
    np.savez_compressed(os.path.join(HERE, 'ref_rankers.npz'), **out)
    print('ref_rankers.npz', len(out), 'arrays')


def small_problem(seed=5, obs_dim=17, act_dim=6, hidden=(64, 64), T=40, table_len=200_003):
    dims = orc.layer_dims(obs_dim, hidden, act_dim)
    P = orc.n_params(dims)
    table = np.random.RandomState(seed).randn(table_len).astype(np.float32)
    theta = (np.random.RandomState(seed + 1).randn(P) * 0.1).astype(np.float32)
    env = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
    return dims, P, table, theta, env


def oracle_vectors():
    out = {}
    dims, P, table, theta, env = small_problem()
    obmean = np.random.RandomState(3).randn(env.obs_dim) * 0.1
    obstd = 0.5 + np.random.RandomState(4).rand(env.obs_dim)
    idx = 12345
    noise = orc.table_get(table, idx, P)
    out['pheno_pos'] = orc.pheno_params(theta, 0.02, noise)
    out['pheno_neg'] = orc.pheno_params(theta, 0.02, -noise)
    layers = orc.unflatten(out['pheno_pos'], dims)
    rews, behv, obs, step = orc.run_model(env, layers, obmean, obstd, 5.0, env.T, batched=False)
    out['rollout_rews'] = np.array(rews)
    out['rollout_fit'] = np.array(orc.reward_result(rews))
    out['rollout_pos'] = np.array(behv[-3:])
    out['rollout_step'] = np.array(step)
    out['obsn'] = orc.normalise_obs(env.obs_stream[:env.T], obmean, obstd, 5.0)
    out['obmean'], out['obstd'] = obmean, obstd
    # two whole generations, 2 virtual ranks x 6 pairs, one coin per evaluation, Adam
    flat = theta.copy()
    opt = orc.AdamOracle(P, 0.01)
    states = [np.random.RandomState(1000), np.random.RandomState(1001)]
    for g in range(2):
        res = orc.generation(table, flat, opt, 0.02, dims, env, [1000, 1001], 6, obmean, obstd, 5.0, env.T, 500, 0.005,
                             coins_per_eval=1, rank_states=states, batched=False)
        out[f'gen{g}_pos'], out[f'gen{g}_neg'], out[f'gen{g}_inds'] = res['pos'], res['neg'], res['inds']
        out[f'gen{g}_w'], out[f'gen{g}_theta'] = res['weights'], flat.copy()
        out[f'gen{g}_steps'] = np.array(res['steps'])
    # NSRA generation (2 objectives)
    flat = theta.copy()
    opt = orc.AdamOracle(P, 0.01)
    archive = np.random.RandomState(17).randn(16, 2)
    res = orc.generation(table, flat, opt, 0.02, dims, env, [1000, 1001], 6, obmean, obstd, 5.0, env.T, 500, 0.005,
                         moo_w=0.5, archive=archive, nov_k=10, coins_per_eval=1, batched=False)
    out['nsra_pos'], out['nsra_neg'], out['nsra_w'], out['nsra_theta'] = res['pos'], res['neg'], res['weights'], flat.copy()
    np.savez_compressed(os.path.join(HERE, 'oracle_vectors.npz'), **out)
    print('oracle_vectors.npz', len(out), 'arrays')


if __name__ == '__main__':
    ref_vectors()
    ref_ranker_vectors()
    oracle_vectors()
