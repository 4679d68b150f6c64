"""Dev: the tcgen05 rollouts (ES_ROLLOUT_TC, ES_ROLLOUT_TC3) against the float32 CUDA-core rollout and a float64 numpy truth.

    python tools/dev_tc2_check.py [quick|full|time]

Prints, per shape: max / rms fitness difference to the float32 kernel, and for the Humanoid shape the error of EACH device
path against float64 arithmetic on the same inputs (what "float32-equivalent" means here: the tensor-core path's error
against the truth is of the size of the float32 path's own error)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from es_pytorch_b200 import _lib                              # noqa: E402
from es_pytorch_b200.engine import get_engine                 # noqa: E402


def truth_f64(table, theta, idx, sigma, sizes, obsn, rew):
    """float64 forward of theta +- sigma*eps on every step, fitness = sum_t <a_t, c_t>"""
    obs, h1, h2, act = sizes
    out = np.zeros((2, len(idx)))
    x = obsn.astype(np.float64)
    c = rew.astype(np.float64)
    for k, i in enumerate(idx):
        eps = table[i:i + len(theta)].astype(np.float64)
        for s, sign in enumerate((1.0, -1.0)):
            w = theta.astype(np.float64) + sign * np.float64(np.float32(sigma)) * eps
            at = 0
            a = x
            for fi, fo in ((obs, h1), (h1, h2), (h2, act)):
                W = w[at:at + fi * fo].reshape(fo, fi); at += fi * fo
                b = w[at:at + fo]; at += fo
                a = np.tanh(a @ W.T + b)
            out[s, k] = (a * c).sum()
    return out


def run_case(eng, obs, act, T, n_pairs, modes, sigma=0.02, seed=0, truth_pairs=0):
    rs = np.random.RandomState(seed + obs + T)
    sizes = [obs, 64, 64, act]
    P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
    L = P + 2_000_000
    table_h = rs.randn(L).astype(np.float32)
    theta_h = (rs.randn(P) * 0.1).astype(np.float32)
    idx_h = rs.randint(0, L - P - 1, size=n_pairs).astype(np.int64)
    obsn_h = np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)
    rew_h = rs.randn(T, act).astype(np.float32)
    table, theta, idx = eng.to_device(table_h), eng.to_device(theta_h), eng.to_device(idx_h)
    obsn, rew = eng.to_device(obsn_h), eng.to_device(rew_h)
    res = {}
    for name, mode in modes:
        fit = torch.zeros(2, n_pairs, dtype=torch.float64, device=eng.device)
        behv = torch.zeros(2, n_pairs, 3, dtype=torch.float32, device=eng.device)
        eng.rollout(table, idx, theta, sigma, sizes, obsn, rew, 0.05, fit[0], fit[1], 1, behv[0], behv[1], mode)
        eng.sync()
        res[name] = (fit.cpu().numpy(), behv.cpu().numpy())
    f32 = res['f32'][0]
    spread = f32.std()
    line = f'obs={obs} act={act} T={T} pairs={n_pairs}: spread {spread:.4g}'
    for name, _ in modes:
        if name == 'f32':
            continue
        d = res[name][0] - f32
        db = np.abs(res[name][1] - res['f32'][1]).max()
        line += f' | {name}-f32: max {np.abs(d).max():.3e} rms {np.sqrt((d ** 2).mean()):.3e} ({np.sqrt((d ** 2).mean()) / spread:.2e} of spread) behv {db:.2e}'
    print(line, flush=True)
    if truth_pairs:
        tr = truth_f64(table_h, theta_h, idx_h[:truth_pairs], sigma, sizes, obsn_h, rew_h)
        for name, _ in modes:
            d = res[name][0][:, :truth_pairs] - tr
            print(f'    {name} vs float64 truth ({truth_pairs} pairs): max {np.abs(d).max():.3e} rms {np.sqrt((d ** 2).mean()):.3e} '
                  f'mean {d.mean():+.2e} ({np.sqrt((d ** 2).mean()) / spread:.2e} of spread)', flush=True)
    return res


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'quick'
    eng = get_engine(0)
    modes = [('f32', _lib.ES_ROLLOUT_F32), ('tc3', _lib.ES_ROLLOUT_TC3), ('tc', _lib.ES_ROLLOUT_TC)]
    if os.environ.get('ES_DEV_MODES'):
        keep = os.environ['ES_DEV_MODES'].split(',')
        modes = [m for m in modes if m[0] in keep or m[0] == 'f32']
    if what in ('quick', 'full'):
        run_case(eng, 24, 6, 160, 8, modes)
        run_case(eng, 24, 6, 1000, 300, modes, truth_pairs=8)
        run_case(eng, 17, 6, 300, 333, modes)                 # no 16-byte aligned rows: builders convert the float32 slice
        run_case(eng, 64, 3, 128, 4, modes)
        run_case(eng, 376, 17, 1000, 200, modes, truth_pairs=32)
    if what == 'full':
        run_case(eng, 5, 1, 130, 3, modes)
        run_case(eng, 63, 32, 129, 5, modes)
        run_case(eng, 17, 6, 100, 400, modes)
        run_case(eng, 376, 17, 1000, 2000, modes, truth_pairs=64)
    if what in ('time', 'full', 'quick'):
        obs, act, T, n = 376, 17, 1000, 10000
        rs = np.random.RandomState(1)
        sizes = [obs, 64, 64, act]
        P = sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))
        g = torch.Generator(device=eng.device).manual_seed(123)
        table = torch.randn(250_000_000, generator=g, device=eng.device, dtype=torch.float32)
        theta = eng.to_device((rs.randn(P) * 0.1).astype(np.float32))
        idx = eng.to_device(rs.randint(0, 250_000_000 - P - 1, size=n).astype(np.int64))
        obsn = eng.to_device(np.clip(rs.randn(T, obs), -5, 5).astype(np.float32))
        rew = eng.to_device(rs.randn(T, act).astype(np.float32))
        fit = torch.zeros(2, n, dtype=torch.float64, device=eng.device)
        for name, mode in modes:
            if name == 'f32':
                continue
            for _ in range(2):
                eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], 1, None, None, mode)
            eng.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], 1, None, None, mode)
            b.record()
            eng.sync()
            print(f'time {name}: {a.elapsed_time(b) / 5:.3f} ms per K=10000 rollout call (incl. prep kernels)', flush=True)


if __name__ == '__main__':
    t0 = time.time()
    main()
    print(f'done in {time.time() - t0:.1f} s')
