#!/usr/bin/env python
"""Time one workload's align kernels (inputs resident in HBM) for several builds of the library.

    python tools/kernel_bench.py c3 [lib.so ...]        # default: the in-tree library

Each library runs in a process of its own (a process loads one libtanw).  Prints GCUPS and ms per
launch, and checks the first 64 pairs against the C oracle.  Used to compare kernel variants
built by `__graft_entry__.build_library(out, defines=[...])` within one gpurun call.
KB_PAIRS=<n> overrides the number of pairs (e.g. 1250: one GPU's share of config 2 over eight)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(workload, lib, steps=8):
    import numpy as np
    import bench
    from text_alignment_b200 import _native
    if lib:
        _native.load(lib)
    npairs = int(os.environ.get('KB_PAIRS', 0)) or bench.WORKLOADS[workload]['default_pairs']
    packed, pairs = bench.make_workload(workload, 0, npairs, 16)
    ctx = _native.Context(0)
    sc = ctx.make_scoring(*bench.DEFAULT_PARAMS)
    ok = True
    if workload != 'c5':
        from oracle import nw_oracle
        k = min(64, npairs)
        sub = bench.pack_pairs(pairs[:k])
        got = ctx.align_batch(*sub, sc)
        osc, _ = nw_oracle.make_scoring(list(bench.DEFAULT_PARAMS[:6]), boundary_gap=bench.DEFAULT_PARAMS[6])
        want = nw_oracle.align_batch_codes(*sub, osc, threads=8)
        ok = np.array_equal(got[2], want[2]) and all(
            np.array_equal(got[0][got[1][i]:got[1][i] + got[2][i]], want[0][want[1][i]:want[1][i] + want[2][i]]) for i in range(k))
    ctx.prepare(*packed, sc)
    for _ in range(3):
        ctx.run()
    ctx.sync()
    best = 1e9
    for _ in range(steps):
        ctx.run()
        ctx.sync()
        best = min(best, ctx.timing()['kernel_ms'])
    cells = int((packed[2].astype(np.int64) * packed[4]).sum())
    print('%-44s %s  %8.1f GCUPS  %8.3f ms  parity %s' % (os.path.basename(lib or 'libtanw.so'), workload,
                                                       cells / best / 1e6, best, 'ok' if ok else 'FAILED'), flush=True)


if __name__ == '__main__':
    if len(sys.argv) >= 3 and sys.argv[1] == '--one':
        one(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != '-' else None)
    else:
        wl = sys.argv[1]
        for lib in (sys.argv[2:] or ['-']):
            subprocess.call([sys.executable, os.path.abspath(__file__), '--one', wl, lib])
