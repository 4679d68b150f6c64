import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dev_e2e_profile.py')).read().split("for _ in range(3): api_generation()")[0])
ts = []
for g in range(60):
    t0 = time.perf_counter(); api_generation(); ts.append((time.perf_counter() - t0) * 1e3)
print('per-generation ms:', ' '.join(f'{x:.1f}' for x in ts))
import gc
print('gc counts', gc.get_count(), 'REG', len(__import__('es_pytorch_b200.devcache', fromlist=['x'])._REG))
# split one generation into phases (wall clock, each ends with a device sync)
def phase(label, fn):
    t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f'{label:28s} {(time.perf_counter()-t0)*1e3:7.3f} ms'); return r
for rep in range(2):
    go = ObStat(env.observation_space.shape, 0)
    gen = phase('_device_generation', lambda: es._device_generation(fit_fn, policy, nt, streams))
    phase('gen.evaluate', lambda: gen.evaluate(1250))
    res = phase('es.test_params (whole)', lambda: es.test_params(comm, 1250, policy, nt, go, fit_fn, streams[0]))
    phase('update_obstat', lambda: policy.update_obstat(go))
    phase('rank', lambda: ranker.rank(res[0], res[1], res[2]))
    phase('approx_grad', lambda: es.approx_grad(policy, ranker, nt, policy.flat_params, 500, 0.005))
