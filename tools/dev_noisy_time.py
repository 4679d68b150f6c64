"""Dev tool: one generation WITH action noise (ac_std = 0.01) at the bench size: time of es_draw_noisy (sequential per
virtual-rank stream) and of the whole generation, for R streams per GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200 import _lib
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200.generation import DeviceGeneration
from es_pytorch_b200.gym.synthetic_env import SyntheticEnv
from es_pytorch_b200.nn.optimizers import Adam

K = int(os.environ.get('K', 10000))
eng = get_engine(0)
obs, act, T = 376, 17, 1000
sizes = [obs, 64, 64, act]; P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
g = torch.Generator(device=eng.device).manual_seed(1)
table = torch.randn(50_000_000, generator=g, device=eng.device)
env = SyntheticEnv(obs, act, T)
obs_dev, rew_dev = env.device_arrays(eng)
theta0 = (np.random.RandomState(7).randn(P) * 0.1).astype(np.float32)
for R in [int(x) for x in os.environ.get('STREAMS', '8,64').split(',')]:
    for ac_std in (0.0, 0.01):
        gen = DeviceGeneration(table, eng.to_device(theta0.copy()), sizes, obs_dev, rew_dev,
                               [np.random.RandomState(1000 + r) for r in range(R)], 0.02, 0.005, Adam(P, 0.01), coins_per_eval=1,
                               save_obs_chance=0.01, rollout_mode=_lib.ES_ROLLOUT_TC3, engine=eng, ac_std=ac_std)
        gen.run(K // R); torch.cuda.synchronize()
        gen.enable_timers(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            gen.run(K // R)
        b.record(); torch.cuda.synchronize()
        kern = {k: float(np.mean([x.elapsed_time(y) for x, y in v])) for k, v in gen.timers.items()}
        print(f'R={R} ac_std={ac_std}: {a.elapsed_time(b) / 3:.3f} ms per generation of K={K}; draw {kern["draw_indices"]:.3f} ms, rollout {kern["rollout"]:.3f} ms', flush=True)
        del gen
