import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dev_e2e_profile.py')).read().split("for _ in range(3): api_generation()")[0])
import es_pytorch_b200.engine as E
# wrap every Engine method + a few host functions with wall-clock timing
import functools, collections
acc = collections.OrderedDict()
def wrap_obj(obj, names, prefix):
    for nme in names:
        orig = getattr(obj, nme)
        def f(*a, __o=orig, __n=prefix + nme, **k):
            t0 = time.perf_counter(); r = __o(*a, **k); acc[__n] = acc.get(__n, 0.0) + (time.perf_counter() - t0) * 1e3; return r
        setattr(obj, nme, f)
wrap_obj(eng, ['draw_indices', 'normalise_obs', 'rollout', 'obs_colsum', 'obstat_accumulate_coins', 'centered_rank',
               'grad_reconstruct', 'adam_step', 'upload_async', 'download_async', 'sync', 'to_device', 'to_host'], 'eng.')
for _ in range(3): api_generation()
gen = fit_fn._gen
wrap_obj(gen, ['load_states', 'store_states', 'set_obstat', 'evaluate'], 'gen.')
for g in range(40):
    acc.clear()
    t0 = time.perf_counter(); api_generation(); dt = (time.perf_counter() - t0) * 1e3
    if dt > 20 or g == 5:
        print(f'gen {g}: {dt:.1f} ms  ' + ' '.join(f'{k}={v:.2f}' for k, v in acc.items() if v > 0.05))
