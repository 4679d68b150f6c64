"""Dev tool: cProfile of the reference-facing API path (es.test_params -> rank -> approx_grad)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200 import dist
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
from es_pytorch_b200 import _lib
eng = get_engine(0); comm = dist.world()
env = SyntheticEnv(376, 17, 1000)
P = 29393
g = torch.Generator(device=eng.device).manual_seed(123)
table = torch.randn(250_000_000, generator=g, device=eng.device)
net = FeedForward([64, 64], torch.nn.Tanh(), env, 0.0, 5)
policy = Policy(net, 0.02, Adam(P, 0.01)); nt = NoiseTable(P, table)
streams = [np.random.RandomState(1000 + r) for r in range(8)]
fit_fn = BatchedRollout(env, 1000, coins_per_eval=1, save_obs_chance=0.01, rank_streams=streams, rollout_mode=_lib.ES_ROLLOUT_TC)
fit_fn.stream_env_from_host = True
ranker = CenteredRanker()
def api_generation():
    gen_obstat = ObStat(env.observation_space.shape, 0)
    pos, neg, inds, steps = es.test_params(comm, 1250, policy, nt, gen_obstat, fit_fn, streams[0])
    policy.update_obstat(gen_obstat)
    ranker.rank(pos, neg, inds)
    es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005)
for _ in range(3): api_generation()
torch.cuda.synchronize(); t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): api_generation()
torch.cuda.synchronize(); pr.disable()
print('ms per generation', (time.perf_counter() - t0) * 100)
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

def timed(label, fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f'{label:40s} {(time.perf_counter()-t0)/n*1e3:8.3f} ms')
gen = fit_fn._gen
timed('gen.evaluate', lambda: gen.evaluate(1250))
timed('gen.evaluate + update', lambda: gen.run(1250))
go = ObStat(env.observation_space.shape, 0)
timed('es.test_params', lambda: es.test_params(comm, 1250, policy, nt, go, fit_fn, streams[0]))
pos, neg, inds, steps = es.test_params(comm, 1250, policy, nt, go, fit_fn, streams[0])
timed('ranker.rank', lambda: ranker.rank(pos, neg, inds))
timed('approx_grad', lambda: es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005))
timed('_device_generation only', lambda: es._device_generation(fit_fn, policy, nt, streams))
timed('load_states', lambda: gen.load_states(streams))
timed('upload obs', lambda: eng.upload_async(gen.obs_stream, env.obs_stream[:gen.T + 1], ('obs', id(gen))))
gen.enable_timers(True)
for _ in range(5): gen.run(1250)
torch.cuda.synchronize()
print({k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in gen.timers.items()})
gen.timers = {}
for _ in range(5): gen.evaluate(1250)
torch.cuda.synchronize()
print('evaluate only', {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in gen.timers.items()})
print('theta finite', bool(torch.isfinite(gen.theta).all()), 'fit finite', bool(torch.isfinite(gen.fit_local).all()), float(gen.theta.abs().max()))
# per-call GPU timing of every engine method inside evaluate-only and inside the API path
import functools
recs = {}
def wrap(name):
    orig = getattr(eng, name)
    @functools.wraps(orig)
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(*a, **k); e1.record()
        recs.setdefault(name, []).append((e0, e1)); return r
    setattr(eng, name, f)
for nme in ['draw_indices', 'normalise_obs', 'rollout', 'obs_colsum', 'obstat_accumulate_coins', 'centered_rank', 'grad_reconstruct', 'adam_step', 'upload_async', 'download_async']:
    wrap(nme)
gen.timers = None
def report(label):
    torch.cuda.synchronize()
    print(label, {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in recs.items()}); recs.clear()
for _ in range(5): gen.evaluate(1250)
report('evaluate-only')
t0 = time.perf_counter()
for _ in range(5): api_generation()
torch.cuda.synchronize(); print('api ms', (time.perf_counter() - t0) / 5 * 1e3)
report('api')
for rep in range(4):
    t0 = time.perf_counter()
    for _ in range(15): api_generation()
    torch.cuda.synchronize(); print('api ms', (time.perf_counter() - t0) / 15 * 1e3)
    report('api rep %d' % rep)
    f = gen.fit_local.flatten()
    print('  fitness min/median/max', float(f.min()), float(f.median()), float(f.max()), 'theta absmax', float(gen.theta.abs().max()))
