import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from es_pytorch_b200.engine import get_engine
eng = get_engine(0)
a = np.random.randn(29393).astype(np.float32)
dst = torch.zeros(29393, device=eng.device)
big = torch.zeros(1 << 28, device=eng.device)
def t(label, fn, n=200):
    t0 = time.perf_counter()
    for _ in range(n): fn()
    print(f'{label:40s} {(time.perf_counter()-t0)/n*1e6:9.1f} us')
ta = torch.from_numpy(a)
t('from_numpy', lambda: torch.from_numpy(np.ascontiguousarray(a)))
t('is_pinned (pageable)', lambda: ta.is_pinned())
pin = torch.empty(29393, pin_memory=True)
t('is_pinned (pinned)', lambda: pin.is_pinned())
t('pinned.copy_(pageable)', lambda: pin.copy_(ta))
t('dst.copy_(pinned, nb)', lambda: dst.copy_(pin, non_blocking=True))
ev = torch.cuda.Event(); ev.record()
t('ev.query', lambda: ev.query())
t('ev.record', lambda: ev.record())
t('upload_async', lambda: eng.upload_async(dst, a, 'k'))
torch.cuda.synchronize()
# with a busy GPU: queue 50 ms of work then time host-side calls
def busy():
    for _ in range(20): big.mul_(1.0001)
busy(); t('busy: is_pinned (pageable)', lambda: ta.is_pinned(), 20)
torch.cuda.synchronize(); busy(); t('busy: dst.copy_(pinned, nb)', lambda: dst.copy_(pin, non_blocking=True), 20)
torch.cuda.synchronize(); busy(); t('busy: upload_async', lambda: eng.upload_async(dst, a, 'k'), 20)
torch.cuda.synchronize()
