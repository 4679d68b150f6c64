"""Dev tool: time the tensor-core rollout at the bench size (K pairs, Humanoid-shaped) with CUDA events and report its
error against the float32 rollout on a sample.  ES_B200_LIB selects a kernel variant (es_pytorch_b200.build.build_variant)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200 import _lib

K = int(os.environ.get('K', 10000))
eng = get_engine(0)
rs = np.random.RandomState(0)
obs, act, T = 376, 17, 1000
sizes = [obs, 64, 64, act]; P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
L = 250_000_000
g = torch.Generator(device=eng.device).manual_seed(1)
table = torch.randn(L, generator=g, device=eng.device)
theta = eng.to_device((rs.randn(P) * 0.1).astype(np.float32))
idx = torch.randint(0, L - P, (K,), generator=g, device=eng.device, dtype=torch.int64)
obsn = eng.to_device(np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)); rew = eng.to_device(rs.randn(T, act).astype(np.float32))
fit = torch.zeros(2, K, dtype=torch.float64, device=eng.device)
times = []
for it in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], mode=_lib.ES_ROLLOUT_TC)
    b.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
n = 512
ref = torch.zeros(2, n, dtype=torch.float64, device=eng.device)
eng.rollout(table, idx[:n], theta, 0.02, sizes, obsn, rew, 0.05, ref[0], ref[1], mode=_lib.ES_ROLLOUT_F32)
f32, tc = ref.cpu().numpy(), fit[:, :n].cpu().numpy()
d = tc - f32
r32 = np.argsort(np.argsort(f32.ravel())); rtc = np.argsort(np.argsort(tc.ravel()))
print(f'lib={os.path.basename(_lib.LIB_PATH)} K={K} ms: min {min(times[2:]):.3f} med {np.median(times[2:]):.3f} | '
      f'rms err/spread {np.sqrt((d ** 2).mean()) / f32.std():.5f} diff-err/diff-std '
      f'{np.sqrt((((tc[0] - tc[1]) - (f32[0] - f32[1])) ** 2).mean()) / (f32[0] - f32[1]).std():.5f} '
      f'rank corr {np.corrcoef(r32, rtc)[0, 1]:.6f}')
