"""Dev tool: device time of the single-policy float32 rollout (es.step's noiseless evaluation, run_model compat path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200 import _lib
eng = get_engine(0)
rs = np.random.RandomState(0)
for obs, act, T in ((376, 17, 1000), (17, 6, 1000)):
    sizes = [obs, 64, 64, act]; P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
    table = eng.to_device(rs.randn(P + 10).astype(np.float32)); theta = eng.to_device((rs.randn(P) * 0.1).astype(np.float32))
    obsn = eng.to_device(np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)); rew = eng.to_device(rs.randn(T, act).astype(np.float32))
    for n in (1, 4, 74, 200):
        idx = torch.zeros(n, dtype=torch.int64, device=eng.device)
        fit = torch.zeros(2, n, dtype=torch.float64, device=eng.device)
        ts = []
        for it in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.rollout(table, idx, theta, 0.0, sizes, obsn, rew, 0.05, fit[0], fit[1], mode=_lib.ES_ROLLOUT_F32); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f'obs={obs} T={T} pairs={n}: {min(ts[1:])*1e3:.1f} us')
