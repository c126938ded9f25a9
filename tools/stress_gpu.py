"""Randomised stress of every kernel family against the C oracle (seeded; prints a summary)."""
import os, sys, random, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from text_alignment_b200 import _native
from oracle import nw_oracle

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
    for c in use:
        got = c.align_batch(*b, c.make_scoring(*params))
        assert np.array_equal(got[2], want[2]), (rd, shape, params)
        assert np.array_equal(got[3].astype(np.float64), wscore), (rd, shape, params)
        for k in range(len(pairs)):
            assert np.array_equal(got[0][got[1][k]:got[1][k]+got[2][k]], want[0][want[1][k]:want[1][k]+want[2][k]]), (rd, shape, params, k, b[2][k], b[4][k])
    total += len(pairs) * len(use)
    print('round', rd, shape, len(pairs), 'pairs', params, 'ok', flush=True)
print('stress ok: %d alignments, %.1f s' % (total, time.time() - t0))
