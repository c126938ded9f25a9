#!/usr/bin/env python
"""Time one workload end to end (tanw_align_batch with pinned host buffers: H2D + tables + kernels +
D2H per call) for a build of the library -- the counterpart of tools/kernel_bench.py.

    python tools/e2e_bench.py c3 [lib.so]        # default: the in-tree library

Used with the TANW_TUNING build to compare chunking choices (TANW_LINE_CHUNKS=k) in one gpurun call."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import bench
    from text_alignment_b200 import _native
    workload = sys.argv[1]
    if len(sys.argv) > 2:
        _native.load(sys.argv[2])
    npairs = int(os.environ.get('KB_PAIRS', 0)) or bench.WORKLOADS[workload]['default_pairs']
    packed, pairs = bench.make_workload(workload, 0, npairs, 16)
    ctx = _native.Context(0)
    sc = ctx.make_scoring(*bench.DEFAULT_PARAMS)

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    sym, t_off, n, o_off, m = [pin(a) for a in packed]
    layout = _native.Context.canonical_ops_layout(n, m)
    P = n.size
    out = (pin(np.zeros(layout[1] + 1, np.uint8)), pin(np.zeros(P, np.int32)), pin(np.zeros((P, 3), np.int32)))
    for _ in range(3):
        ctx.align_batch(sym, t_off, n, o_off, m, sc, out=out, layout=layout)
    steps = 10
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.align_batch(sym, t_off, n, o_off, m, sc, out=out, layout=layout)
    dt = (time.perf_counter() - t0) / steps
    cells = float((n.astype(np.int64) * m).sum())
    tm = ctx.timing()
    print('%-28s %s  e2e %8.1f GCUPS  %8.3f ms  chunks %d' % (os.path.basename(sys.argv[2]) if len(sys.argv) > 2 else 'libtanw.so',
                                                             workload, cells / dt / 1e9, dt * 1e3, getattr(tm, 'chunks', -1)))


if __name__ == '__main__':
    main()
