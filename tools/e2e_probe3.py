"""Spikes in the e2e leg: per-call wall time and library host timers, with and without NVML sampling."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from text_alignment_b200.textSeqCompare import get_context
which = sys.argv[1] if len(sys.argv) > 1 else 'c4'
npairs = bench.WORKLOADS[which]['default_pairs']
packed, pairs = bench.make_workload(which, 0, npairs, 8)
buf, t_off, n, o_off, m = packed
ctx = get_context(0)
sc = ctx.make_scoring(*bench.DEFAULT_PARAMS)
tb = torch.empty(buf.size, dtype=torch.uint8, pin_memory=True); pb = tb.numpy(); pb[...] = buf
ops_cap = int((n.astype(np.int64) + m).sum())
t_ops = torch.empty(ops_cap, dtype=torch.uint8, pin_memory=True)
t_len = torch.empty(n.size, dtype=torch.int32, pin_memory=True)
t_sc = torch.empty((n.size, 3), dtype=torch.int32, pin_memory=True)
out = (t_ops.numpy(), t_len.numpy(), t_sc.numpy())
def loop(tag, k=24):
    rows = []
    for it in range(k):
        t0 = time.perf_counter(); ctx.align_batch(pb, t_off, n, o_off, m, sc, out=out); w = (time.perf_counter()-t0)*1e3
        tm = ctx.timing()
        rows.append((w, tm['host_prepare_ms'], tm['host_run_ms'], tm['host_fetch_ms'], tm['kernel_ms']))
    print(tag)
    for r in rows: print('   wall %.2f  prepare %.2f run %.2f fetch %.2f | kernel %.2f' % r)
loop('no sampler')
smp = bench.ClockSampler(0); smp.start(); time.sleep(0.2)
loop('nvml sampler 50 ms')
print(smp.stop(0, 1e18))
