import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dev_e2e_profile.py')).read().split("for _ in range(3): api_generation()")[0])
recs = {}
def wrap(name):
    orig = getattr(eng, name)
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(*a, **k); e1.record()
        recs.setdefault(name, []).append((e0, e1)); return r
    setattr(eng, name, f)
for nme in ['draw_indices', 'normalise_obs', 'rollout', 'obs_colsum', 'obstat_accumulate_coins', 'centered_rank', 'grad_reconstruct', 'adam_step']:
    wrap(nme)
for _ in range(3): api_generation()
for g in range(24):
    recs.clear()
    t0 = time.perf_counter(); api_generation(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    if dt > 20 or g == 1:
        print(f'gen {g}: {dt:.1f} ms', {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in recs.items()},
              'mt_pos', fit_fn._gen.mt_pos.cpu().numpy().tolist())
