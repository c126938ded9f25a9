"""Bisect the host overhead of the e2e leg: torch import, pinned buffers, NVML sampler."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from text_alignment_b200.textSeqCompare import get_context
packed, pairs = bench.make_workload('c2', 0, 10000, 8)
buf, t_off, n, o_off, m = packed
ctx = get_context(0)
sc = ctx.make_scoring(*bench.DEFAULT_PARAMS)
def loop(tag, b, out=None, k=5):
    ts = []
    for it in range(k):
        t0 = time.perf_counter(); ctx.align_batch(b, t_off, n, o_off, m, sc, out=out); ts.append((time.perf_counter()-t0)*1e3)
    print(tag, ' '.join('%.1f' % t for t in ts), ctx.timing())
loop('pageable, no torch   ', buf)
import torch
torch.cuda.set_device(0)
loop('pageable, torch      ', buf)
tb = torch.empty(buf.size, dtype=torch.uint8, pin_memory=True); pb = tb.numpy(); pb[...] = buf
ops_cap = int((n.astype(np.int64) + m).sum())
t_ops = torch.empty(ops_cap, dtype=torch.uint8, pin_memory=True)
t_len = torch.empty(n.size, dtype=torch.int32, pin_memory=True)
t_sc = torch.empty((n.size, 3), dtype=torch.int32, pin_memory=True)
out = (t_ops.numpy(), t_len.numpy(), t_sc.numpy())
loop('pinned in            ', pb)
loop('pinned in+out        ', pb, out)
smp = bench.ClockSampler(0); smp.start(); time.sleep(0.2)
loop('pinned + nvml sampler', pb, out)
print(smp.stop(0, 1e18))
for it in range(3):
    t0 = time.perf_counter(); ctx.prepare(pb, t_off, n, o_off, m, sc)
    t1 = time.perf_counter(); ctx.run(); ctx.sync()
    t3 = time.perf_counter(); ctx.fetch()
    t4 = time.perf_counter()
    print('phases: prepare %.2f run+sync %.2f fetch(pageable out) %.2f' % ((t1-t0)*1e3, (t3-t1)*1e3, (t4-t3)*1e3))
