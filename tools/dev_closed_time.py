"""Dev tool: one generation on the CLOSED-LOOP synthetic env at the bench size (Humanoid-shaped, K pairs, T = 1000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200.generation import DeviceGeneration
from es_pytorch_b200.gym.synthetic_env import ClosedLoopEnv
from es_pytorch_b200.nn.optimizers import Adam

K, T, R = int(os.environ.get('K', 10000)), int(os.environ.get('T', 1000)), 8
eng = get_engine(0)
obs, act = 376, 17
sizes = [obs, 64, 64, act]; P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
g = torch.Generator(device=eng.device).manual_seed(1)
table = torch.randn(50_000_000, generator=g, device=eng.device)
env = ClosedLoopEnv(obs, act, T)
obs_dev, rew_dev = env.device_arrays(eng)
theta0 = (np.random.RandomState(7).randn(P) * 0.1).astype(np.float32)
gen = DeviceGeneration(table, eng.to_device(theta0.copy()), sizes, obs_dev, rew_dev, [np.random.RandomState(1000 + r) for r in range(R)],
                       0.02, 0.005, Adam(P, 0.01), coins_per_eval=1, save_obs_chance=0.01, engine=eng, closed=env.device_closed(eng))
gen.run(K // R); torch.cuda.synchronize()
gen.enable_timers(True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(2):
    gen.run(K // R)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 2
kern = {k: float(np.mean([x.elapsed_time(y) for x, y in v])) for k, v in gen.timers.items()}
print(f'closed-loop generation K={K} T={T}: {ms:.2f} ms ({K / ms * 1e3:.0f} pairs/s); rollout {kern["rollout"]:.2f} ms', flush=True)
