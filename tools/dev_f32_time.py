"""Dev tool: time the float32 rollout (Humanoid shape) with CUDA events; ES_B200_LIB selects a variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200 import _lib
K = int(os.environ.get('K', 1184))
eng = get_engine(0); rs = np.random.RandomState(0)
obs, act, T = 376, 17, 1000
sizes = [obs, 64, 64, act]; P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
L = 50_000_000
g = torch.Generator(device=eng.device).manual_seed(1)
table = torch.randn(L, generator=g, device=eng.device)
theta = eng.to_device((rs.randn(P) * 0.1).astype(np.float32))
idx = torch.randint(0, L - P, (K,), generator=g, device=eng.device, dtype=torch.int64)
obsn = eng.to_device(np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)); rew = eng.to_device(rs.randn(T, act).astype(np.float32))
fit = torch.zeros(2, K, dtype=torch.float64, device=eng.device)
ts = []
for it in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], mode=_lib.ES_ROLLOUT_F32); b.record()
    torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f'lib={os.path.basename(_lib.LIB_PATH)} K={K}: {min(ts[1:]):.3f} ms  ({min(ts[1:]) * 1e3 / K:.2f} us/pair) checksum {float(fit.sum()):.6f}')
