import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dev_e2e_profile.py')).read().split("for _ in range(3): api_generation()")[0])
def series(label, n=30):
    ts = []
    for g in range(n):
        t0 = time.perf_counter(); api_generation(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f'{label:34s}', ' '.join(f'{x:.0f}' for x in ts))
for _ in range(3): api_generation()
series('baseline')
fit_fn.stream_env_from_host = False
series('no env re-upload')
fit_fn.stream_env_from_host = True
fit_fn.save_obs_chance = 0.0; fit_fn._gen.save_obs_chance = 0.0
series('save_obs_chance 0')
fit_fn.coins_per_eval = 0; fit_fn._gen = None
series('no coins (new gen)')
