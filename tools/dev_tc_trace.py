"""Dev tool: cycle-stamp timeline of CTA 0 of the tensor-core rollout (ES_TC_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
from es_pytorch_b200 import _lib
eng = get_engine(0)
trace = torch.zeros(4 * 512, dtype=torch.int64, device=eng.device)
os.environ['ES_TC_TRACE'] = hex(trace.data_ptr())
rs = np.random.RandomState(0)
obs, act, T, n_pairs = 376, 17, 1000, 148 * 3
sizes = [obs, 64, 64, act]; P = sum(i*o+o for i, o in zip(sizes[:-1], sizes[1:]))
L = 50_000_000
g = torch.Generator(device=eng.device).manual_seed(1)
table = torch.randn(L, generator=g, device=eng.device)
theta = eng.to_device((rs.randn(P)*0.1).astype(np.float32))
idx = torch.randint(0, L-P, (n_pairs,), generator=g, device=eng.device, dtype=torch.int64)
obsn = eng.to_device(np.clip(rs.randn(T, obs), -5, 5).astype(np.float32)); rew = eng.to_device(rs.randn(T, act).astype(np.float32))
fit = torch.zeros(2, n_pairs, dtype=torch.float64, device=eng.device)
for _ in range(2):
    trace.zero_()
    eng.rollout(table, idx, theta, 0.02, sizes, obsn, rew, 0.05, fit[0], fit[1], mode=_lib.ES_ROLLOUT_TC)
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(4, 512)
t0 = t[t > 0].min()
mma, epi = t[0], t[1]
print('tile | MMA: L1 first-chunk, L1 done-issue, L2+ issue, L3+ issue | EPI: wait D1, got D1, H1N arrived, got D2+, got D2-, wait D3+, got D3+, got D3-')
for g_ in range(24):
    m = [(mma[4*g_+k]-t0) if mma[4*g_+k] else -1 for k in range(4)]
    e = [(epi[8*g_+k]-t0) if epi[8*g_+k] else -1 for k in range(8)]
    print(f'{g_:3d} | {m} | {e}')


print('EPI detail (relative to got D1): epi1 first tmem_ld done, epi1+ done | rel. to got D2+: ld done, epi2+ done | rel. to got D3+: ld done, epi3+ done | rel. to got D3-: ld done, epi3- done')
for g_ in range(0, 24, 2):
    x = t[2][8*g_:8*g_+8]; e = epi[8*g_:8*g_+8]
    print(g_, [int(x[0]-e[1]), int(x[1]-e[1])], [int(x[2]-e[3]), int(x[3]-e[3])], [int(x[4]-e[6]), int(x[5]-e[6])], [int(x[6]-e[7]), int(x[7]-e[7])])
print('builder warp 0 per image j: [start, slot free, L1 rows done, image ready]')
for j in range(3): print(j, [int(t[3][4*j+k]-t0) if t[3][4*j+k] else -1 for k in range(4)])
