"""Randomised stress of every kernel family against the C oracle (seeded; prints a summary).

    python tools/stress_gpu.py [rounds] [seed] [--lib path/to/libtanw_checked.so]

With --lib the run uses another build of the library -- the TANW_CHECKED build
(`__graft_entry__.build_library(out, defines=['TANW_CHECKED'])`), whose device assertions (arena
bounds of every pointer slot, op-string lengths, stamped hand-over records, bounded spins, the
16-bit range rule) turn into an AssertionError here; compute-sanitizer is closed on this pool."""
import os, sys, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from text_alignment_b200 import _native
from oracle import nw_oracle
if '--lib' in sys.argv:
    at = sys.argv.index('--lib')
    _native.load(sys.argv[at + 1])
    del sys.argv[at:at + 2]
    print('library:', _native._lib._name)

def pack(pairs):
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode('latin-1'), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32); m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
    return buf, t_off, n, t_off + n, m

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 123)
ctx = _native.Context(0)
long_ctx = _native.Context(0); long_ctx.set_long_threshold(1)
band_ctx = _native.Context(0); band_ctx.set_long_threshold(1)       # chained stripes cut into row bands
total = 0
t0 = time.time()
for rd in range(rounds):
    alpha = rng.choice(['ab', 'abc', 'acgt', 'abcdefghilmnopqrstuvxy .'])
    shape = rng.choice(['lines', 'pages', 'mixed', 'tall', 'wide'])
    pairs = []
    for _ in range(rng.randint(1, 400)):
        if shape == 'lines':   n, m = rng.randint(0, 130), rng.randint(0, 140)
        elif shape == 'pages': n, m = rng.randint(100, 700), rng.randint(100, 2600)
        elif shape == 'tall':  n, m = rng.randint(500, 5000), rng.randint(1, 60)
        elif shape == 'wide':  n, m = rng.randint(1, 40), rng.randint(500, 4000)
        else:                  n, m = rng.choice([(rng.randint(0, 130), rng.randint(0, 140)), (rng.randint(100, 500), rng.randint(100, 1500))])
        t = ''.join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.6 and n:
            o = list(t)
            for _ in range(rng.randint(0, max(1, n // 5))):
                o[rng.randrange(len(o))] = rng.choice(alpha)
            o = ''.join(o)[:m] + ''.join(rng.choice(alpha) for _ in range(max(0, m - n)))
        else:
            o = ''.join(rng.choice(alpha) for _ in range(m))
        pairs.append((t, o))
    if rng.random() < 0.5:
        params = (8, -4, -7, -7, -3, 0, -1)
    else:
        params = (rng.randint(0, 12), rng.randint(-10, 3), rng.randint(-10, 2), rng.randint(-10, 2),
                  rng.randint(-6, 2), rng.randint(-6, 2), rng.randint(-5, 2))
    b = pack(pairs)
    sc, _ = nw_oracle.make_scoring(list(params[:6]), boundary_gap=params[6])
    want = nw_oracle.align_batch_codes(*b, sc, threads=16)
    wscore = np.where(want[3] <= -1e99, -1073741824, want[3])
    use = [ctx] + ([long_ctx] if len(pairs) <= 60 else [])
    if len(pairs) <= 25 and shape != 'tall':
        band_ctx.set_long_band_rows(rng.choice([3, 17, 40, 64, 129, 300]))
        use.append(band_ctx)
    ctx.set_line_kernel(rng.choice([1, 1, 2, 0]))
    for c in use:
        got = c.align_batch(*b, c.make_scoring(*params))
        assert np.array_equal(got[2], want[2]), (rd, shape, params)
        assert np.array_equal(got[3].astype(np.float64), wscore), (rd, shape, params)
        for k in range(len(pairs)):
            assert np.array_equal(got[0][got[1][k]:got[1][k]+got[2][k]], want[0][want[1][k]:want[1][k]+want[2][k]]), (rd, shape, params, k, b[2][k], b[4][k])
    total += len(pairs) * len(use)
    print('round', rd, shape, len(pairs), 'pairs', params, 'ok', flush=True)
# a big batch of lines: the chunked pipeline of tanw_align_batch, and per-pair scoring systems
ctx.set_line_kernel(1)
pairs = []
for _ in range(60000):
    n, m = rng.randint(1, 125), rng.randint(1, 130)
    t = ''.join(rng.choice('abcdefghil ') for _ in range(n))
    pairs.append((t, t[:m] + ''.join(rng.choice('abcdefghil ') for _ in range(max(0, m - n)))))
b = pack(pairs)
sc, _ = nw_oracle.make_scoring([8, -4, -7, -7, -3, 0], boundary_gap=-1)
want = nw_oracle.align_batch_codes(*b, sc, threads=16)
got = ctx.align_batch(*b, ctx.make_scoring(8, -4, -7, -7, -3, 0, -1))
assert ctx.timing()['chunks'] > 1 and np.array_equal(got[2], want[2])
for k in range(0, len(pairs), 7):
    assert np.array_equal(got[0][got[1][k]:got[1][k]+got[2][k]], want[0][want[1][k]:want[1][k]+want[2][k]]), k
total += len(pairs)
systems = [(8, -4, -7, -7, -3, 0, -1), (5, -4, -2, -7, 0, -5, -1), (7, 2, 3, -4, -1, 1, -1)]
sub = pack(pairs[:300])
idx = np.array([rng.randrange(3) for _ in range(300)], dtype=np.int32)
got = ctx.align_batch_multi(*sub, systems, idx)
for s_i, prm in enumerate(systems):
    sc, _ = nw_oracle.make_scoring(list(prm[:6]), boundary_gap=prm[6])
    want = nw_oracle.align_batch_codes(*sub, sc, threads=16)
    for k in np.nonzero(idx == s_i)[0]:
        assert np.array_equal(got[0][got[1][k]:got[1][k]+got[2][k]], want[0][want[1][k]:want[1][k]+want[2][k]]), (s_i, k)
total += 300
print('stress ok: %d alignments, %.1f s' % (total, time.time() - t0))
