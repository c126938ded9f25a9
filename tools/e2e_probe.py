"""Where does the end-to-end time of tanw_align_batch go?  Times prepare / run / fetch."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from text_alignment_b200.textSeqCompare import get_context

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
packed, pairs = bench.make_workload('c2', 0, npairs, 8)
buf, t_off, n, o_off, m = packed
ctx = get_context(0)
sc = ctx.make_scoring(*bench.DEFAULT_PARAMS)
for it in range(6):
    t0 = time.perf_counter(); ctx.prepare(buf, t_off, n, o_off, m, sc)
    t1 = time.perf_counter(); ctx.run()
    t2 = time.perf_counter(); ctx.sync()
    t3 = time.perf_counter(); out = ctx.fetch()
    t4 = time.perf_counter()
    tm = ctx.timing()
    print('iter %d prepare %.2f launch %.2f kernel-wait %.2f fetch %.2f ms | lib h2d %.2f k %.2f d2h %.2f' % (
        it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, tm['h2d_ms'], tm['kernel_ms'], tm['d2h_ms']))
for it in range(4):
    t0 = time.perf_counter(); ctx.align_batch(buf, t_off, n, o_off, m, sc); t1 = time.perf_counter()
    print('align_batch %.2f ms' % ((t1-t0)*1e3))
