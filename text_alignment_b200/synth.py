"""Seeded synthetic page / line pairs.

There is no network and the reference ships no data, so every benchmark and
parity test runs on these: Latin-like transcripts built from a fixed syllable
list (alphabet ``[a-z ]``, every word has a vowel, no regex metacharacters and
no ``_``/``~``) and an OCR string derived from the transcript by per-character
deletion / insertion / substitution noise plus inserted runs until a target
length is reached.  The digests of the three Appendix-C vectors are pinned in
``tests/test_synth.py`` so generator drift shows up separately from aligner
drift.
"""
import random

SYL = ['a', 'e', 'o', 'do', 'mi', 'nus', 'al', 'le', 'lu', 'ia', 'glo', 'ri', 'in', 'ex', 'cel',
       'sis', 'de', 'et', 'ter', 'ra', 'pax', 'ho', 'ni', 'bus', 'bo', 'ne', 'vo', 'lun', 'ta',
       'tis', 'sanc', 'tus', 'qui', 'tol', 'lis', 'pec', 'ca', 'mun', 'di', 'no', 'bis', 're',
       'gem', 'cu', 'ius', 'quo', 'ni', 'am', 've', 'ni', 'te']
OCR_ALPHA = 'abcdefghilmnopqrstuvxy .'


def gen_transcript(rng, nchars):
    words = []
    total = 0
    while total < nchars:
        w = ''.join(rng.choice(SYL) for _ in range(rng.randint(1, 4)))
        words.append(w)
        total += len(w) + 1
    return ' '.join(words)[:nchars].rstrip(' ')


def gen_ocr(rng, t, sub, indel, target_len, run_lo, run_hi):
    out = []
    for c in t:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(rng.choice(OCR_ALPHA))
        out.append(rng.choice(OCR_ALPHA) if rng.random() < sub else c)
    while len(out) < target_len:
        p = rng.randrange(len(out) + 1)
        k = min(rng.randint(run_lo, run_hi), target_len - len(out))
        out[p:p] = [rng.choice(OCR_ALPHA) for _ in range(k)]
    return ''.join(out)


def make_pair(seed, n, m, run_lo, run_hi, sub=0.20, indel=0.05):
    rng = random.Random(seed)
    t = gen_transcript(rng, n)
    o = gen_ocr(rng, t, sub, indel, m, run_lo, run_hi)
    return t, o


# ---- the five BASELINE.json configs --------------------------------------

def c1_page():
    """Config 1: single Salzinnes-shaped page, seed 1001, n=1200, m=1500."""
    return make_pair(1001, 1200, 1500, 5, 40)


def c2_pair(k):
    """Config 2, pair k: seed 2000000+k, n~U[1000,1600] drawn first, m=round(1.25 n)."""
    seed = 2000000 + k
    rng = random.Random(seed)
    n = rng.randint(1000, 1600)
    m = int(round(1.25 * n))
    t = gen_transcript(rng, n)
    o = gen_ocr(rng, t, 0.20, 0.05, m, 5, 40)
    return t, o


def c3_pair(k):
    """Config 3, line pair k: seed 3000000+k, n,m~U[40,120], runs 2-6."""
    seed = 3000000 + k
    rng = random.Random(seed)
    n = rng.randint(40, 120)
    m = rng.randint(40, 120)
    t = gen_transcript(rng, n)
    o = gen_ocr(rng, t, 0.20, 0.05, m, 2, 6)
    return t, o


def c4_pair(k):
    """Config 4, St. Gall-shaped page k: seed 4000000+k, n~U[600,1000], m=n*U[2,4], runs 50-400."""
    seed = 4000000 + k
    rng = random.Random(seed)
    n = rng.randint(600, 1000)
    m = int(n * rng.uniform(2.0, 4.0))
    t = gen_transcript(rng, n)
    o = gen_ocr(rng, t, 0.20, 0.05, m, 50, 400)
    return t, o


def c5_pair(n=80000, m=100000):
    """Config 5: whole-manuscript pair, seed 5001, n=80 000 (transcript), m=100 000 (OCR)."""
    return make_pair(5001, n, m, 5, 400)


# ---- fast numpy generator for bulk throughput runs -------------------------------------------
# The Python generator above costs ~1 ms per page; 10k pages is fine, 1M line pairs is not.
# bench.py uses the exact generators for parity samples and this vectorised one (same shape
# statistics, different random stream) where only throughput is measured; it says which.

def bulk_pairs_numpy(seed, count, n_lo, n_hi, m_of_n, sub=0.20, indel=0.05):
    """Return (codes uint8 concatenated, t_off, n, o_off, m) for `count` pairs.

    Transcript characters are drawn from the Appendix-C syllable inventory letter
    frequencies with spaces every ~6 chars; OCR = transcript with substitution noise,
    random deletions, and random insertions up to length m_of_n(n).
    Codes are ASCII bytes (so they decode with bytes.decode('ascii')).
    """
    import numpy as np
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(''.join(SYL).encode(), dtype=np.uint8)
    alpha = np.frombuffer(OCR_ALPHA.encode(), dtype=np.uint8)
    ns = rng.integers(n_lo, n_hi + 1, size=count)
    ms = np.asarray([int(m_of_n(int(n), rng)) for n in ns], dtype=np.int64)
    t_off = np.zeros(count, dtype=np.int64)
    o_off = np.zeros(count, dtype=np.int64)
    total = int(ns.sum() + ms.sum())
    buf = np.empty(total, dtype=np.uint8)
    pos = 0
    for k in range(count):
        n = int(ns[k]); m = int(ms[k])
        t = letters[rng.integers(0, len(letters), size=n)]
        sp = rng.random(n) < (1.0 / 6.0)
        t = np.where(sp, np.uint8(32), t)
        if n:
            t[0] = letters[0]; t[-1] = letters[1]
        keep = rng.random(n) >= indel / 2
        o = t[keep].copy()
        subm = rng.random(o.size) < sub
        o[subm] = alpha[rng.integers(0, len(alpha), size=int(subm.sum()))]
        if o.size > m:
            o = o[:m]
        while o.size < m:
            p = int(rng.integers(0, o.size + 1))
            r = min(int(rng.integers(2, 41)), m - o.size)
            o = np.concatenate([o[:p], alpha[rng.integers(0, len(alpha), size=r)], o[p:]])
        t_off[k] = pos; buf[pos:pos + n] = t; pos += n
        o_off[k] = pos; buf[pos:pos + m] = o; pos += m
    return buf, t_off, ns.astype(np.int64), o_off, ms


# ---- synthetic OCR character boxes for the consumer ------------------------------

def gen_char_boxes(rng, ocr, line_lo=40, line_hi=60, w_lo=10, w_hi=30, x0=100, y0=200, pitch=140, height=60):
    """(char, ul, lr) per OCR character: lines of 40-60 chars, widths U[10,30] px, fixed pitch."""
    boxes = []
    x, y, left = x0, y0, rng.randint(line_lo, line_hi)
    for c in ocr:
        w = rng.randint(w_lo, w_hi)
        boxes.append((c, (x, y), (x + w, y + height)))
        x += w
        left -= 1
        if left == 0:
            x, y, left = x0, y + pitch, rng.randint(line_lo, line_hi)
    return boxes


def drop_vowelless_tail(t):
    """Truncation can leave a last word fragment without a vowel, which the reference's
    syllabifier cannot handle (latinSyllabification.py:71); drop it (SURVEY App. C note)."""
    words = t.split(' ')
    while words and not any(v in words[-1] for v in 'aeiouy'):
        words.pop()
    return ' '.join(words)


def make_page(seed, n, m, run_lo=5, run_hi=40, abbreviations=False):
    """Transcript + OCR character boxes of one synthetic page."""
    rng = random.Random(seed)
    t = drop_vowelless_tail(gen_transcript(rng, n))
    o = gen_ocr(rng, t, 0.20, 0.05, m, run_lo, run_hi)
    if abbreviations:
        o = o.replace('dominus', 'dns', 2).replace('alleluia', 'alla', 1)
        if 'um ' in o:
            o = o.replace('um ', u'ū ', 1)
    return t, gen_char_boxes(rng, o)
