"""Time per row of the chained-stripe fill for 1, 2, 4, ... stripes (traceback skipped):
separates a stripe's own speed from the cost of the hand-over between stripes.
Run with TANW_DEBUG_SKIP_TRACE=1 TANW_LONG_C=8."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from text_alignment_b200 import _native

ctx = _native.Context(0)
ctx.set_long_threshold(1)
rng = np.random.default_rng(5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
for m in (256, 512, 1024, 4096, 16384, 65536):
    t = rng.integers(97, 101, size=n, dtype=np.uint8)
    o = rng.integers(97, 101, size=m, dtype=np.uint8)
    buf = np.concatenate([t, o])
    args = (buf, np.array([0], np.int64), np.array([n], np.int32), np.array([n], np.int64), np.array([m], np.int32))
    sc = ctx.make_scoring(8, -4, -7, -7, -3, 0, -1)
    ctx.prepare(*args, sc)
    best = 1e9
    for _ in range(4):
        ctx.run(); ctx.sync()
        best = min(best, ctx.timing()['kernel_ms'])
    stripes = (m + 255) // 256
    skew = 46 * stripes
    print('m=%6d stripes=%4d  %.3f ms  %.1f ns/row  (%.1f ns per row+skew step)' %
          (m, stripes, best, best * 1e6 / n, best * 1e6 / (n + skew)), flush=True)
