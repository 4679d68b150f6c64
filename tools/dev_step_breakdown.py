"""Dev tool: where the host time of es.step's fused route goes (perf_counter around its phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200 import _lib, dist, devcache
from es_pytorch_b200.core import es
from es_pytorch_b200.core.noisetable import NoiseTable
from es_pytorch_b200.core.policy import Policy
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200.gym.batched import BatchedRollout
from es_pytorch_b200.gym.synthetic_env import SyntheticEnv
from es_pytorch_b200.nn.nn import FeedForward
from es_pytorch_b200.nn.obstat import ObStat
from es_pytorch_b200.nn.optimizers import Adam
from es_pytorch_b200.utils.rankers import CenteredRanker
from es_pytorch_b200.utils.reporters import Reporter
eng = get_engine(0)
obs, act, T, K, R = 376, 17, 1000, 10000, 8
sizes = [obs, 64, 64, act]; P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
g = torch.Generator(device=eng.device).manual_seed(123)
table = torch.randn(250_000_000, generator=g, device=eng.device)
env = SyntheticEnv(obs, act, T)
net = FeedForward([64, 64], torch.nn.Tanh(), env, 0.0, 5)
policy = Policy(net, 0.02, Adam(P, 0.01)); nt = NoiseTable(P, table)
streams = [np.random.RandomState(1000 + r) for r in range(R)]
fit_fn = BatchedRollout(env, T, coins_per_eval=1, save_obs_chance=0.01, rank_streams=streams, rollout_mode=getattr(_lib, os.environ.get('MODE', 'ES_ROLLOUT_TC3')))
ranker = CenteredRanker(); comm = dist.world()
class C(dict): __getattr__ = dict.__getitem__
cfg = C(general=C(policies_per_gen=2 * K // R, batch_size=500), policy=C(l2coeff=0.005))
rep = Reporter()
seg = {}
def tick(name, t0):
    t = time.perf_counter(); seg.setdefault(name, []).append(t - t0); return t
orig_dev_gen, orig_sync = es._device_generation, eng.sync
def dev_gen(*a, **k):
    t0 = time.perf_counter(); r = orig_dev_gen(*a, **k); tick('pre: _device_generation', t0); return r
def sync():
    t0 = time.perf_counter(); orig_sync(); tick('sync', t0)
es._device_generation = dev_gen; eng.sync = sync
for it in range(25):
    t0 = time.perf_counter()
    tr, ob = es.step(cfg, comm, policy, nt, env, fit_fn, streams[0], ranker, rep)
    t1 = tick('es.step total', t0)
    policy.update_obstat(ob)
    tick('update_obstat', t1)
for k, v in seg.items(): print(f'{k:28s} {1e3 * np.median(v[5:]):.3f} ms')
tot = np.median(seg['es.step total'][5:]); pre = np.median(seg['pre: _device_generation'][5:]); sy = np.median(seg['sync'][5:])
print(f'launch+post (total - pre - sync) = {1e3 * (tot - pre - sy):.3f} ms')

# ---- phase marks inside _step_fused ----
es.TRACE = {}
for it in range(30):
    tr, ob = es.step(cfg, comm, policy, nt, env, fit_fn, streams[0], ranker, rep)
    policy.update_obstat(ob)
T = {k: np.array(v[5:]) for k, v in es.TRACE.items()}
es.TRACE = None
names = list(T)
for a, b in zip(names[:-1], names[1:]):
    print(f'{a:22s} -> {b:22s} {1e3 * np.median(T[b] - T[a]):.3f} ms')

# ---- host profile of the same loop (cProfile): which python frames the non-GPU time is spent in ----
import cProfile, pstats, io
pr = cProfile.Profile()
pr.enable()
for it in range(40):
    tr, ob = es.step(cfg, comm, policy, nt, env, fit_fn, streams[0], ranker, rep)
    policy.update_obstat(ob)
pr.disable()
sio = io.StringIO()
pstats.Stats(pr, stream=sio).sort_stats('tottime').print_stats(28)
print(sio.getvalue()[:6000])
