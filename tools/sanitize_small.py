"""Small end-to-end run for compute-sanitizer: every kernel variant on tiny inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from text_alignment_b200 import _native, synth, textSeqCompare as tsc
from oracle import nw_oracle

def pack(pairs):
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode('latin-1'), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32); m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
    return buf, t_off, n, t_off + n, m

def check(ctx, pairs, params, subst=None):
    b = pack(pairs)
    got = ctx.align_batch(*b, ctx.make_scoring(*params, subst=subst))
    if subst is None:
        sc, _ = nw_oracle.make_scoring(list(params[:6]), boundary_gap=params[6])
        want = nw_oracle.align_batch_codes(*b, sc, threads=4)
        assert np.array_equal(got[2], want[2])
        for k in range(len(pairs)):
            assert np.array_equal(got[0][got[1][k]:got[1][k]+got[2][k]], want[0][want[1][k]:want[1][k]+want[2][k]]), k
    return got

ctx = _native.Context(0)
sizes = [(0, 0), (0, 5), (5, 0), (1, 1), (3, 130), (40, 33), (70, 100), (100, 128), (33, 513), (64, 600), (200, 1100), (5000, 3)]
pairs = [synth.make_pair(10 + k, n, m, 1, 9) if n and m else ('a' * n, 'b' * m) for k, (n, m) in enumerate(sizes)]
check(ctx, pairs, (8, -4, -7, -7, -3, 0, -1))            # EYZ kernels (lines + pages)
check(ctx, pairs, (7, -3, -4, -9, -1, -2, -2))           # general kernels
K = 128
tab = (np.arange(K * K, dtype=np.int32).reshape(K, K) % 7) - 3
check(ctx, pairs, (0, 0, -5, -6, -1, -2, -1), subst=tab)  # substitution-table kernels
ctx.set_long_threshold(1)                                # chained-pass path
check(ctx, pairs[3:11], (8, -4, -7, -7, -3, 0, -1))
check(ctx, pairs[3:11], (7, -3, -4, -9, -1, -2, -2))
print('int32 peak', ctx.measure_int32_peak(0) > 0)
ctx.close()
print('sanitize_small ok')
