"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): two processes, one GPU each, shard the
generation (fitness allgather + one grad allreduce over NCCL) and must reproduce the oracle's 2-rank generation."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = '''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from oracle import es_oracle as orc
from es_pytorch_b200 import dist, _lib
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200.generation import DeviceGeneration
from es_pytorch_b200.nn.optimizers import Adam
comm = dist.init_from_env('nccl')
eng = get_engine(int(os.environ['LOCAL_RANK']))
obs_dim, act_dim, hidden, T, n = 17, 6, (64, 64), 48, 10
dims = orc.layer_dims(obs_dim, hidden, act_dim); P = orc.n_params(dims)
rs = np.random.RandomState(0)
table = rs.randn(P + 200_000).astype(np.float32); theta = (rs.randn(P) * 0.1).astype(np.float32)
env = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
seeds = [1000 + 2 * r for r in range(comm.size * 2)]            # 2 virtual ranks per GPU
mine = seeds[2 * comm.rank: 2 * comm.rank + 2]
gen = DeviceGeneration(eng.to_device(table), eng.to_device(theta.copy()), [obs_dim, *hidden, act_dim],
                       eng.to_device(env.obs_stream), eng.to_device(env.rew_vec), [np.random.RandomState(s) for s in mine],
                       0.02, 0.005, Adam(P, 0.01), coins_per_eval=1, save_obs_chance=0.2, comm=comm, engine=eng)
flat, opt = theta.copy(), orc.AdamOracle(P, 0.01)
states = [np.random.RandomState(s) for s in seeds]
for g in range(2):
    gen.run(n)
    torch.cuda.synchronize()
    ref = orc.generation(table, flat, opt, 0.02, dims, env, seeds, n, np.zeros(obs_dim), np.ones(obs_dim), 5.0, T, 500, 0.005,
                         coins_per_eval=1, rank_states=states)
    k0 = gen.k_begin
    assert np.array_equal(gen.idx.cpu().numpy(), ref['inds'][k0:k0 + gen.k_local].astype(np.int64)), 'indices'
    assert np.abs(gen.fpos_all.cpu().numpy() - ref['pos']).max() < 1e-3, 'fitness allgather order'
    assert np.array_equal(gen.weights.cpu().numpy(), ref['weights'][k0:k0 + gen.k_local]), 'shard weights'
    assert np.abs(gen.theta.cpu().numpy() - flat).max() < 2e-6, 'theta after allreduce'
# every rank holds the identical theta (replicated optimizer step on an allreduced gradient)
th = gen.theta.clone(); allth = torch.empty(comm.size, P, device=eng.device)
comm.allgather_into(allth, th)
assert torch.equal(allth[0], allth[comm.size - 1])
os.write(1, ('MULTI_OK_%d\\n' % comm.rank).encode())
'''


_STEP_WORKER = '''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
from oracle import es_oracle as orc
from es_pytorch_b200 import dist
from es_pytorch_b200.core import es
from es_pytorch_b200.core.noisetable import NoiseTable
from es_pytorch_b200.core.policy import Policy
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200.gym.batched import BatchedRollout
from es_pytorch_b200.gym.synthetic_env import SyntheticEnv
from es_pytorch_b200.nn.nn import FeedForward
from es_pytorch_b200.nn.optimizers import Adam
from es_pytorch_b200.utils.rankers import CenteredRanker
from es_pytorch_b200.utils.reporters import Reporter
comm = dist.init_from_env('nccl')
eng = get_engine(int(os.environ['LOCAL_RANK']))
obs_dim, act_dim, hidden, T, n = 17, 6, (64, 64), 48, 10
dims = orc.layer_dims(obs_dim, hidden, act_dim); P = orc.n_params(dims)
rs = np.random.RandomState(0)
table = rs.randn(P + 200_000).astype(np.float32); theta = (rs.randn(P) * 0.1).astype(np.float32)
spec = orc.SyntheticEnvSpec(obs_dim, act_dim, T)
env = SyntheticEnv(obs_dim, act_dim, T)
net = FeedForward(list(hidden), torch.nn.Tanh(), env, 0.0, 5)
policy = Policy(net, 0.02, Adam(P, 0.01)); policy.flat_params[...] = theta
nt = NoiseTable(P, table)
seeds = [1000 + 2 * r for r in range(comm.size * 2)]            # 2 virtual ranks per GPU
streams = [np.random.RandomState(s) for s in seeds[2 * comm.rank: 2 * comm.rank + 2]]
fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=0.2, rank_streams=streams)
class C(dict): __getattr__ = dict.__getitem__
cfg = C(general=C(policies_per_gen=2 * n * comm.size, batch_size=500), policy=C(l2coeff=0.005))
ranker = CenteredRanker()
assert es._can_fuse_step(comm, policy, fit_fn, ranker)
flat, opt = theta.copy(), orc.AdamOracle(P, 0.01)
states = [np.random.RandomState(s) for s in seeds]
for g in range(2):
    tr, ob = es.step(cfg, comm, policy, nt, env, fit_fn, streams[0], ranker, Reporter())
    # es.step = the generation + the noiseless evaluation, whose fit_fn call draws one more coin on EVERY rank (es.py:48)
    ref = orc.es_step(table, flat, opt, 0.02, dims, spec, states, n, np.zeros(obs_dim), np.ones(obs_dim), 5.0, T, 500, 0.005,
                      coins_per_eval=1)
    assert np.array_equal(ranker.noise_inds, ref['inds']), 'all ranks see all indices, rank-major'
    assert np.array_equal(ranker.ranked_fits, ref['weights']) and ranker.n_fits_ranked == ref['n_ranked'], 'weights'
    assert np.abs(ranker.fits_pos - ref['pos']).max() < 1e-3
    assert np.abs(policy.flat_params - flat).max() < 2e-6, 'theta'
    layers = orc.unflatten(flat, dims)
    rews, b, _, _ = orc.run_model(spec, layers, np.zeros(obs_dim), np.ones(obs_dim), 5.0, T, batched=True)
    assert abs(tr.result[0] - orc.reward_result(rews)[0]) <= 1e-5 * max(1.0, np.abs(rews).sum()), 'noiseless result'
for s, st in zip(streams, states[2 * comm.rank: 2 * comm.rank + 2]):
    assert np.array_equal(s.get_state()[1], st.get_state()[1]) and s.get_state()[2] == st.get_state()[2]
os.write(1, ('STEP_OK_%d\\n' % comm.rank).encode())
'''


def _run_two(script):
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    return subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                           '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                          capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_es_step_fused(tmp_path):
    """es.step (single-synchronisation route) on two processes == the oracle's 4-rank generation."""
    script = tmp_path / 'ws.py'
    script.write_text(_STEP_WORKER.format(root=ROOT))
    out = _run_two(script)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert 'STEP_OK_0' in out.stdout and 'STEP_OK_1' in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_generation(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER.format(root=ROOT))
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert 'MULTI_OK_0' in out.stdout and 'MULTI_OK_1' in out.stdout
