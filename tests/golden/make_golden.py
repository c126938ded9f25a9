"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/textSeqCompare.py) in the build container.

    python tests/golden/make_golden.py            # ~1-2 min (two full pages at ~18 s each)

The reference cannot travel to the GPU box, so its outputs are committed as small JSON
fixtures.  Each record carries the inputs, the scoring system, the aligned sequences as an
op string ('0' diag, '1' transcript char vs '_', '2' '_' vs OCR char), the end-corner scores
(M, X, Y)[n][m] (NEG = -1e100 -> null) and a sha256 of the packed pointer matrix
(PM | PX<<2 | PY<<4)[1:,1:] captured from the reference's own np.zeros arrays.
"""
import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_loader          # noqa: E402
from text_alignment_b200 import synth  # noqa: E402
import scorers                         # noqa: E402


def ops_from_alignment(tra, ocr):
    out = []
    for a, b in zip(tra, ocr):
        if b == '_' and a != '_':
            out.append('1')
        elif a == '_' and b != '_':
            out.append('2')
        else:
            out.append('0')
    return ''.join(out)


def run_ref(T, O, system):
    """T, O lists; system a reference-style scoring list (callable given by name)."""
    sysobj = system
    if system is not None and isinstance(system[0], str):
        sysobj = [scorers.SCORERS[system[0]]] + list(system[1:])
    tra, ocr, mats = ref_loader.reference_align_full(T, O, sysobj)
    n, m = len(T), len(O)
    # op string is unambiguous only if the inputs contain no '_' (true for all fixtures)
    ops = ops_from_alignment(tra, ocr)
    assert len(tra) == len(ocr)
    ptr = (mats['PM'].astype(np.uint8) | (mats['PX'].astype(np.uint8) << 2) |
           (mats['PY'].astype(np.uint8) << 4))[1:, 1:]
    end = [mats[k][n][m] for k in ('M', 'X', 'Y')]
    end = [None if v <= -1e99 else (int(v) if float(v).is_integer() else float(v)) for v in end]
    return dict(ops=ops, end=end, ptr_sha256=hashlib.sha256(np.ascontiguousarray(ptr).tobytes()).hexdigest(),
                tra=tra, ocr=ocr)


def main():
    # ---- 1. known-answer tests -------------------------------------
    kat_pairs = [('abc', 'abc'), ('abc', ''), ('', 'abc'), ('', ''), ('a', 'b'), ('ab', 'ba'),
                 ('dominus', 'dns'), ('dns', 'dominus'), ('alleluia', 'a l l e l u y a'),
                 ('gloria', 'xxxxgloriaxxxx'), ('xxxxgloriaxxxx', 'gloria'),
                 ('ca', 'aa'), ('a', 'a'), ('a', 'c')]
    kats = []
    for t, o in kat_pairs:
        r = run_ref(list(t), list(o), None)
        kats.append(dict(T=t, O=o, system=None, ops=r['ops'], end=r['end'], ptr_sha256=r['ptr_sha256'],
                         tra=''.join(r['tra']), ocr=''.join(r['ocr'])))
    for t, o, s in [('aaaa', 'aa', [10, -5, -7, -7]), ('abcabc', 'abc', [5, -4, -2, -7, 0, -5])]:
        r = run_ref(list(t), list(o), s)
        kats.append(dict(T=t, O=o, system=s, ops=r['ops'], end=r['end'], ptr_sha256=r['ptr_sha256'],
                         tra=''.join(r['tra']), ocr=''.join(r['ocr'])))
    # the reference's own demo (textSeqCompare.py:180-190): 2-char elements, 4-parameter system
    seq1 = 'Lorem ipsum dolor sit amet, consectetur adipiscing elit '
    seq2 = 'LoLorem fipsudolor ..... sit eamet, c.nnr adizisdcing eelitellit'
    e1 = [seq1[2 * x] + seq1[2 * x + 1] for x in range(len(seq1) // 2)]
    e2 = [seq2[2 * x] + seq2[2 * x + 1] for x in range(len(seq2) // 2)]
    r = run_ref(e1, e2, [10, -5, -7, -7])
    demo = dict(T=e1, O=e2, system=[10, -5, -7, -7], ops=r['ops'], end=r['end'], ptr_sha256=r['ptr_sha256'],
                tra='|'.join(r['tra']), ocr='|'.join(r['ocr']))
    r = run_ref(list(seq1), list(seq2), None)
    demo_chars = dict(T=seq1, O=seq2, system=None, ops=r['ops'], end=r['end'], ptr_sha256=r['ptr_sha256'],
                      tra=''.join(r['tra']), ocr=''.join(r['ocr']))
    json.dump(dict(kats=kats, demo=demo, demo_chars=demo_chars), open(os.path.join(HERE, 'kats.json'), 'w'),
              indent=1, sort_keys=True)

    # ---- 2. random differential vectors ------------------------------------------------------
    rng = random.Random(20261018)
    rnd = []
    systems = [None, [10, -5, -7, -7], [5, -4, -2, -7, 0, -5], [8, -4, -7, -7, -3, 0],
               [1, -1, -1, -1], [3, -2, 0, 0, -1, -1], [2, 2, -1, -3, -2, 0], [0, 0, 0, 0],
               [7, -3, -4, -9, -1, -2], [11, -10, -2, -2, -5, -5], [5, -7, -7, -2, 0, -3],
               [1, 3, -2, -2, -1, -1], [4, -4, 2, -3, -1, 1],
               ['vowel_aware', -7, -7, -3, 0], ['confusable', -4, -6, -1, -2], ['asymmetric', -3, -3, -1, -1]]
    for k in range(420):
        alpha = rng.choice(['ab', 'abc', 'acgt', 'aeioudnm s', 'abcdefghilmnopqrstuvxy .'])
        n = rng.randint(0, 34) if k % 7 else rng.randint(0, 3)
        m = rng.randint(0, 34) if k % 5 else rng.randint(0, 3)
        t = ''.join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.5 and n:
            # OCR-like derivative so that long matching diagonals and ties both occur
            o = synth.gen_ocr(rng, t, 0.25, 0.2, m, 1, 4)[:m] if m else ''
        else:
            o = ''.join(rng.choice(alpha) for _ in range(m))
        s = systems[k % len(systems)] if k % 3 else (
            [rng.randint(0, 12), rng.randint(-10, 2)] + [rng.randint(-10, 1) for _ in range(4)])
        r = run_ref(list(t), list(o), s)
        rnd.append(dict(T=t, O=o, system=s, ops=r['ops'], end=r['end'], ptr_sha256=r['ptr_sha256']))
    # medium pairs (a few strips wide for the GPU kernel; several hundred columns)
    for k, (n, m, runs) in enumerate([(150, 200, (2, 8)), (200, 150, (2, 8)), (97, 333, (10, 60)),
                                      (260, 130, (1, 3)), (129, 257, (3, 20)), (64, 1025, (50, 200))]):
        t, o = synth.make_pair(7000 + k, n, m, *runs)
        s = [None, [10, -5, -7, -7], [5, -4, -2, -7, 0, -5], None, ['vowel_aware', -7, -7, -3, 0], None][k]
        r = run_ref(list(t), list(o), s)
        rnd.append(dict(T=t, O=o, system=s, ops=r['ops'], end=r['end'], ptr_sha256=r['ptr_sha256']))
    json.dump(rnd, open(os.path.join(HERE, 'random_pairs.json'), 'w'), indent=0)

    # ---- 3. the Appendix C seeded vectors (inputs regenerated from the seed; digests only) -----
    app = []
    for tag, seed, n, m, lo, hi in [('salzinnes_page', 1001, 1200, 1500, 5, 40),
                                    ('stgall_page', 1002, 800, 2400, 50, 400),
                                    ('line_pair', 1003, 80, 100, 2, 6)]:
        t, o = synth.make_pair(seed, n, m, lo, hi)
        r = run_ref(list(t), list(o), None)
        app.append(dict(tag=tag, seed=seed, n=n, m=m, run_lo=lo, run_hi=hi,
                        input_sha256=hashlib.sha256((t + '\n' + o).encode()).hexdigest(),
                        align_sha256=hashlib.sha256((''.join(r['tra']) + '\n' + ''.join(r['ocr'])).encode()).hexdigest(),
                        ptr_sha256=r['ptr_sha256'], end=r['end'], L=len(r['ops']),
                        gaps_tra=r['tra'].count('_'), gaps_ocr=r['ocr'].count('_'),
                        ops_sha256=hashlib.sha256(r['ops'].encode()).hexdigest()))
        print(tag, app[-1]['end'], app[-1]['L'], app[-1]['align_sha256'])
    json.dump(app, open(os.path.join(HERE, 'appendix_c.json'), 'w'), indent=1, sort_keys=True)


# ---- 4. the consumer: reference process() driven with mocked Gamera / OCR ----

class _Dim(object):
    def __init__(self, c, r):
        self.ncols, self.nrows = c, r


class _Img(object):
    dim = _Dim(3000, 4000)


def run_reference_process(A, transcript, boxes, params=None):
    """Unmodified alignToOCR.process with preprocessing / OCR patched out; angle 0."""
    import tempfile
    chars = [A.CharBox(c, ul, lr) for c, ul, lr in boxes]
    saved = (A.preproc.preprocess_images, A.preproc.identify_text_lines, A.perform_ocr_with_ocropus)
    A.preproc.preprocess_images = lambda raw: (_Img(), None, 0.0)
    A.preproc.identify_text_lines = lambda image, eroded: ([], [100, 240, 380], None)
    A.perform_ocr_with_ocropus = lambda *a, **k: chars
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        return A.process(_Img(), transcript, 'nomodel', seq_align_params=params, wkdir_name='wk')
    finally:
        os.chdir(cwd)
        A.preproc.preprocess_images, A.preproc.identify_text_lines, A.perform_ocr_with_ocropus = saved


def consumer_golden():
    A = ref_loader.load_aligntoocr()
    pages = []
    specs = [(9000 + k, 120 + 15 * k, 150 + 22 * k, 2, 12, k % 2 == 1, None) for k in range(10)]
    specs.append((9100, 260, 900, 50, 200, True, None))                  # St. Gall-like long insertions
    specs.append((9101, 200, 250, 2, 10, False, [5, -4, -2, -7, 0, -5]))
    specs.append((9102, 600, 750, 5, 40, True, None))
    for seed, n, m, lo, hi, abbr, params in specs:
        t, boxes = synth.make_page(seed, n, m, lo, hi, abbreviations=abbr)
        syl_boxes, _, _, all_chars = run_reference_process(A, t, boxes, params)
        pages.append(dict(seed=seed, transcript=t, params=params,
                          boxes=[[c, list(ul), list(lr)] for c, ul, lr in boxes],
                          expanded_ocr=''.join(x.char for x in all_chars),
                          syl_boxes=[[b.char, [int(v) for v in b.ul], [int(v) for v in b.lr]] for b in syl_boxes]))
        print('page', seed, len(t), len(boxes), len(syl_boxes))
    # syllabifier vectors: the reference demo sentence + generated words
    rng = random.Random(99)
    words = ['euouae', 'cuius', 'eius', '', 'quaecumque', 'ejus', 'michi', 'antiphonum', 'assistens',
             'alleluya', 'dixit', 'extra', 'exhibeamus', 'oeix', 'aeiou', 'ththa', 'strophe']
    for _ in range(400):
        w = ''.join(rng.choice(synth.SYL) for _ in range(rng.randint(1, 4)))
        words.append(w)
    for _ in range(200):
        w = ''.join(rng.choice('aeiouybcdfghlmnpqrstvx') for _ in range(rng.randint(1, 8)))
        if any(v in w for v in 'aeiouy'):
            words.append(w)
    # a word whose vowels are all swallowed by consonant clusters ('quqs') never terminates in
    # the reference (latinSyllabification.py:71): record null for those
    import signal

    class _Hang(Exception):
        pass

    def _alarm(signum, frame):
        raise _Hang()
    signal.signal(signal.SIGALRM, _alarm)
    syl = []
    for w in words:
        signal.setitimer(signal.ITIMER_REAL, 0.5)
        try:
            syl.append([w, A.latsyl.syllabify_word(w)])
        except _Hang:
            syl.append([w, None])
        finally:
            signal.setitimer(signal.ITIMER_REAL, 0)
    json.dump(dict(pages=pages, syllables=syl), open(os.path.join(HERE, 'consumer.json'), 'w'), indent=0)


def wide_golden():
    """Pairs with more than 256 distinct elements (16-bit symbol codes on the device).  Elements
    are single characters drawn from a large alphabet, stored as code points."""
    rng = random.Random(20261018)
    out = []
    for n, m, k_alpha, system in [(260, 320, 3000, None), (200, 330, 2500, [5, -4, -2, -7, 0, -5]),
                                  (300, 220, 2000, [10, -5, -7, -7]), (180, 420, 1500, ['asymmetric', -3, -3, -1, -1])]:
        alphabet = [chr(0x100 + c) for c in range(k_alpha)]        # no '_' (gap symbol), no ' '
        T = [rng.choice(alphabet) for _ in range(n)]
        O = list(T)
        for _ in range(n // 4):
            O[rng.randrange(len(O))] = rng.choice(alphabet)
        cut = rng.randrange(0, max(1, len(O) - 20))
        del O[cut:cut + 15]
        while len(O) < m:
            at = rng.randrange(0, len(O) + 1)
            O[at:at] = [rng.choice(alphabet) for _ in range(min(m - len(O), rng.randint(1, 25)))]
        O = O[:m]
        assert len(set(T) | set(O)) > 256
        r = run_ref(T, O, system)
        out.append(dict(T=[ord(c) for c in T], O=[ord(c) for c in O], system=system, ops=r['ops'], end=r['end'],
                        ptr_sha256=r['ptr_sha256']))
        print('wide', n, m, len(set(T) | set(O)), r['end'], len(r['ops']))
    json.dump(out, open(os.path.join(HERE, 'wide_pairs.json'), 'w'), indent=0)


if __name__ == '__main__':
    if '--only-consumer' in sys.argv:
        consumer_golden()
    elif '--only-wide' in sys.argv:
        wide_golden()
    else:
        main()
        consumer_golden()
        wide_golden()
