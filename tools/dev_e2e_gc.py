import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dev_e2e_profile.py')).read().split("for _ in range(3): api_generation()")[0])
for _ in range(3): api_generation()
def series(label):
    ts = []
    for g in range(30):
        t0 = time.perf_counter(); api_generation(); ts.append((time.perf_counter() - t0) * 1e3)
    print(label, ' '.join(f'{x:.1f}' for x in ts))
series('gc on ')
gc.disable(); series('gc off'); gc.enable()
t0 = time.perf_counter(); n = gc.collect(); print('gc.collect', n, (time.perf_counter() - t0) * 1e3, 'ms')
gc.callbacks.append(lambda phase, info: print('   GC', phase, info) if phase == 'stop' else None)
series('gc on2')
