"""Exploratory: tensor-core rollout vs float32 rollout on the device (prints error statistics)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200 import _lib

eng = get_engine(0)
def case(obs, act, T, n_pairs, seed=0):
    rs = np.random.RandomState(seed)
    sizes = [obs, 64, 64, act]
    P = sum(i*o+o for i, o in zip(sizes[:-1], sizes[1:]))
    L = P + 2_000_000
    table = eng.to_device(rs.randn(L).astype(np.float32))
    theta = eng.to_device((rs.randn(P)*0.1).astype(np.float32))
    idx = eng.to_device(rs.randint(0, L-P, size=n_pairs).astype(np.int64))
    obsn = eng.to_device(np.clip(rs.randn(T, obs), -5, 5).astype(np.float32))
    rew = eng.to_device(rs.randn(T, act).astype(np.float32))
    out = {}
    for name, mode in (('f32', _lib.ES_ROLLOUT_F32), ('tc', _lib.ES_ROLLOUT_TC)):
        fit = torch.zeros(2, n_pairs, dtype=torch.float64, device=eng.device)
        behv = torch.zeros(2, n_pairs, 3, dtype=torch.float32, device=eng.device)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], 1, behv[0], behv[1], mode)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out[name] = (fit.cpu().numpy(), behv.cpu().numpy(), dt)
    f32, tc = out['f32'][0], out['tc'][0]
    d = tc - f32
    diff32, difftc = f32[0]-f32[1], tc[0]-tc[1]
    print(f'obs={obs} act={act} T={T} pairs={n_pairs}: f32 {out["f32"][2]*1e3:.2f} ms, tc {out["tc"][2]*1e3:.2f} ms')
    print(f'  fitness std {f32.std():.4f}  max|tc-f32| {np.abs(d).max():.5f}  rms {np.sqrt((d**2).mean()):.5f}')
    print(f'  antithetic diff std {diff32.std():.4f}  rms err of (f+ - f-) {np.sqrt(((difftc-diff32)**2).mean()):.5f}')
    print(f'  behv max diff {np.abs(out["tc"][1]-out["f32"][1]).max():.5f}')
    if n_pairs >= 64:
        r32 = np.argsort(np.argsort(f32.ravel())); rtc = np.argsort(np.argsort(tc.ravel()))
        print(f'  rank corr {np.corrcoef(r32, rtc)[0,1]:.6f}')
case(17, 6, 100, 8)
case(376, 17, 1000, 64, seed=1)
case(17, 6, 1000, 300, seed=2)
case(5, 1, 130, 3, seed=3)
