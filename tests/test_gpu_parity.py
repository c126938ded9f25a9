"""Parity tests proper: the CUDA path (through the C ABI) against the golden vectors of the
unmodified reference and against the CPU oracle on seeded inputs.  Bit-exact: integer
scores, op strings, aligned sequences.  Run on the B200 box with `-m gpu`."""
import hashlib
import io
import random
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import golden_elems, ops_string, resolve_system
from text_alignment_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tsc():
    from text_alignment_b200 import textSeqCompare
    textSeqCompare.get_context(0)          # fails loudly without libtanw.so / a B200
    return textSeqCompare


@pytest.fixture(scope='module')
def oracle():
    from oracle import nw_oracle
    return nw_oracle


def _end(end):
    return tuple(None if (v is None or v <= -1e99) else int(v) for v in end)


def _pack(pairs):
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode('latin-1'), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
    m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    lens = n.astype(np.int64) + m
    t_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64) if len(pairs) else np.zeros(0, np.int64)
    return buf, t_off, n, t_off + n, m


DEFAULT = (8, -4, -7, -7, -3, 0, -1)


def _check_packed_vs_oracle(tsc, oracle, pairs, params=DEFAULT, threads=8):
    buf, t_off, n, o_off, m = _pack(pairs)
    ops, ops_off, ops_len, scores = tsc.align_packed(buf, t_off, n, o_off, m, params)
    sc, _ = oracle.make_scoring(list(params[:6]), boundary_gap=params[6])
    r_ops, r_off, r_len, r_end = oracle.align_batch_codes(buf, t_off, n, o_off, m, sc, threads=threads)
    assert ops_len.tolist() == r_len.tolist()
    assert ops_off.tolist() == r_off.tolist()
    for k in range(len(pairs)):
        a = ops[ops_off[k]:ops_off[k] + ops_len[k]]
        b = r_ops[r_off[k]:r_off[k] + r_len[k]]
        assert np.array_equal(a, b), 'pair %d (n=%d, m=%d): op strings differ' % (k, n[k], m[k])
        want = _end(r_end[k].tolist())
        got = tuple(None if v == -1073741824 else int(v) for v in scores[k].tolist())
        assert got == want, 'pair %d: end scores %s vs %s' % (k, got, want)


def test_kats_against_reference_golden(tsc, kats):
    for rec in kats['kats'] + [kats['demo'], kats['demo_chars']]:
        T, O = golden_elems(rec)
        tra, ocr, score = tsc.perform_alignment(T, O, scoring_system=resolve_system(rec['system']),
                                                return_scores=True)
        sep = '|' if rec is kats['demo'] else ''
        assert sep.join(tra) == rec['tra'] and sep.join(ocr) == rec['ocr'], (rec['T'], rec['O'])
        assert tuple(score) == _end(rec['end'])


def test_random_golden_pairs(tsc, random_pairs):
    """436 vectors produced by the unmodified reference: all three scoring-system forms
    (incl. callables), empty and ragged inputs, tie-heavy small alphabets."""
    by_system = {}
    for rec in random_pairs:
        by_system.setdefault(repr(rec['system']), []).append(rec)
    for recs in by_system.values():
        system = resolve_system(recs[0]['system'])
        pairs = [golden_elems(r) for r in recs]
        out = tsc.perform_alignment_batch(pairs, scoring_system=system, return_scores=True)
        for rec, (tra, ocr, score) in zip(recs, out):
            assert ops_string(tra, ocr) == rec['ops'], (rec['T'], rec['O'], rec['system'])
            assert tuple(score) == _end(rec['end']), (rec['T'], rec['O'], rec['system'])


def test_wide_alphabet_golden_pairs(tsc, wide_pairs):
    """Vectors of the unmodified reference on pairs with more than 256 distinct elements
    (16-bit symbol codes, incl. a callable scorer over 409 symbols)."""
    for rec in wide_pairs:
        tra, ocr, score = tsc.perform_alignment(rec['T'], rec['O'], resolve_system(rec['system']), return_scores=True)
        assert ops_string(tra, ocr) == rec['ops'], rec['system']
        assert tuple(score) == _end(rec['end']), rec['system']


def test_appendix_c_pages(tsc, appendix_c):
    """The three seeded page/line vectors of SURVEY.md Appendix C (digests from the reference)."""
    for rec in appendix_c:
        t, o = synth.make_pair(rec['seed'], rec['n'], rec['m'], rec['run_lo'], rec['run_hi'])
        tra, ocr, score = tsc.perform_alignment(list(t), list(o), return_scores=True)
        got = hashlib.sha256((''.join(tra) + '\n' + ''.join(ocr)).encode()).hexdigest()
        assert got == rec['align_sha256'], rec['tag']
        assert tuple(score) == _end(rec['end'])
        assert len(tra) == rec['L'] and tra.count('_') == rec['gaps_tra'] and ocr.count('_') == rec['gaps_ocr']


def test_quirk_discriminators(tsc):
    """boundary constant, first-index tie-break, start state."""
    assert tsc.perform_alignment(list('ca'), list('aa')) == (list('ca_'), list('_aa'))
    assert tsc.perform_alignment(list('a'), list('a')) == (list('a'), list('a'))
    assert tsc.perform_alignment(list('a'), list('c')) == (list('a'), list('c'))


def test_module_gap_extend_read_at_call_time(tsc, oracle, monkeypatch):
    for g in (-1, -4, 0, 2):
        monkeypatch.setattr(tsc, 'gap_extend', g)
        for T, O in [('ca', 'aa'), ('dominus', 'dns'), ('abcabc', 'abc'), ('xxxxgloriaxxxx', 'gloria')]:
            assert tuple(tsc.perform_alignment(list(T), list(O))) == \
                tuple(oracle.perform_alignment(list(T), list(O), None, boundary_gap=g))


def test_empty_and_degenerate(tsc, oracle):
    pairs = [('', ''), ('a', ''), ('', 'b'), ('abc', ''), ('', 'abc'), ('a', 'a'), ('a', 'b'),
             ('a' * 40, 'a'), ('a', 'a' * 40), ('ab' * 70, 'ba' * 70), ('a' * 129, 'a' * 129),
             ('a' * 33, 'b' * 1025)]
    _check_packed_vs_oracle(tsc, oracle, pairs)
    out = tsc.perform_alignment_batch([(list(t), list(o)) for t, o in pairs])
    assert out[0] == ([], [])
    assert out[3] == (list('abc'), list('___'))


def test_every_strip_width_and_pass_boundary(tsc, oracle):
    """m sweeps every remainder strip width (C = 4..32) and the 1024-column pass boundary."""
    rng = random.Random(11)
    pairs = []
    for m in list(range(1, 40)) + [127, 128, 129, 255, 256, 257, 383, 385, 511, 513, 640, 767, 769, 896,
                                   1000, 1023, 1024, 1025, 1151, 1153, 2047, 2048, 2049, 2200, 3100]:
        n = rng.choice([1, 2, 31, 32, 33, 64, 100])
        t, o = synth.make_pair(100000 + m, n, m, 1, 12)
        pairs.append((t, o))
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_tall_pairs(tsc, oracle):
    pairs = [synth.make_pair(300 + k, n, m, 2, 9) for k, (n, m) in
             enumerate([(2000, 5), (1500, 40), (3000, 130), (1024, 1024), (1025, 1023), (700, 1300)])]
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_random_parameters_small_alphabet(tsc, oracle):
    """Tie-heavy inputs under random integer scoring systems, incl. positive gaps/mismatches."""
    rng = random.Random(2024)
    for trial in range(12):
        alpha = rng.choice(['ab', 'abc', 'acgt'])
        pairs = []
        for _ in range(60):
            n, m = rng.randint(0, 150), rng.randint(0, 150)
            pairs.append((''.join(rng.choice(alpha) for _ in range(n)), ''.join(rng.choice(alpha) for _ in range(m))))
        params = (rng.randint(0, 12), rng.randint(-10, 3), rng.randint(-10, 2), rng.randint(-10, 2),
                  rng.randint(-6, 1), rng.randint(-6, 1), rng.randint(-5, 2))
        _check_packed_vs_oracle(tsc, oracle, pairs, params)


def test_grid_search_parameter_vectors(tsc, oracle):
    """The reference's only batch workload: evaluate_text_alignment.py:181-188 sweeps
    {5,8,11}x{-4,-7,-10}x{-2,-5,-7}^2x{0,-3,-5}^2; sample the grid on fixed pairs."""
    from itertools import product
    grid = list(product([5, 8, 11], [-4, -7, -10], [-2, -5, -7], [-2, -5, -7], [0, -3, -5], [0, -3, -5]))
    rng = random.Random(9)
    pairs = [synth.make_pair(500 + k, 180 + 10 * k, 230 + 7 * k, 3, 25) for k in range(3)]
    for p in rng.sample(grid, 24):
        _check_packed_vs_oracle(tsc, oracle, pairs, tuple(p) + (-1,))


def test_callable_scorer_medium(tsc, oracle):
    import scorers
    t, o = synth.make_pair(8123, 300, 420, 3, 30)
    for name in ('vowel_aware', 'confusable', 'asymmetric'):
        system = [scorers.SCORERS[name], -7, -6, -3, -1]
        got = tsc.perform_alignment(list(t), list(o), scoring_system=system, return_scores=True)
        want = oracle.perform_alignment(list(t), list(o), system, full=True)
        assert (got[0], got[1]) == (want[0], want[1])
        assert tuple(got[2]) == _end(want[2]['end'])


def test_wide_symbol_kernel_matches_byte_kernels(tsc):
    """16-bit symbol codes (tanw_set_symbol_bytes) run on the page kernel only; the same pairs
    with the same codes widened to uint16 must give exactly what the byte kernels (page, line
    and chained-stripe routes) give."""
    from text_alignment_b200 import _native
    pairs = [synth.c2_pair(40 + k) for k in range(5)] + [synth.c3_pair(k) for k in range(60)] + \
            [('', 'abc'), ('abc', ''), ('', ''), ('a', 'b'), synth.make_pair(77, 700, 33, 2, 9)]
    buf, t_off, n, o_off, m = _pack(pairs)
    ctx = _native.Context(0)
    try:
        table = np.fromfunction(lambda a, b: ((a * 7 + b * 3) % 11) - 6, (256, 256)).astype(np.int32)
        for params, subst in [(DEFAULT, None), ((5, -4, -2, -7, 0, -5, -1), None), ((7, 2, -4, 3, -1, 1, 0), None),
                              ((0, 0, -3, -4, -1, -2, -1), table)]:
            narrow = ctx.align_batch(buf, t_off, n, o_off, m, ctx.make_scoring(*params, subst=subst))
            wide = ctx.align_batch(buf.astype(np.uint16), t_off, n, o_off, m, ctx.make_scoring(*params, subst=subst))
            assert narrow[2].tolist() == wide[2].tolist()
            assert np.array_equal(narrow[3], wide[3])
            for k in range(len(pairs)):
                assert np.array_equal(narrow[0][narrow[1][k]:narrow[1][k] + narrow[2][k]],
                                      wide[0][wide[1][k]:wide[1][k] + wide[2][k]]), k
    finally:
        ctx.close()


def test_more_than_256_distinct_elements(tsc):
    """The reference takes any hashable elements (textSeqCompare.py:13-22); pairs with more than
    256 distinct ones get 16-bit codes.  Checked against the pure-Python restatement."""
    from oracle import py_port
    rng = random.Random(12)
    alphabet = [chr(0x400 + k) for k in range(500)]
    T = [rng.choice(alphabet) for _ in range(260)]
    O = list(T)
    for _ in range(60):
        O[rng.randrange(len(O))] = rng.choice(alphabet)
    del O[40:70]
    O[100:100] = [rng.choice(alphabet) for _ in range(35)]
    assert len(set(T) | set(O)) > 256
    for system in (None, [5, -4, -2, -7, 0, -5], [lambda a, b: 6 if a == b else (-1 if ord(a) % 2 == ord(b) % 2 else -5), -7, -6, -3, -1]):
        assert tuple(tsc.perform_alignment(T, O, system)) == tuple(py_port.perform_alignment(T, O, system))
    # non-string elements, mixed with an ordinary pair in one batch
    T2 = [(rng.randrange(400), 'x') for _ in range(150)]
    O2 = T2[:60] + [(rng.randrange(400), 'x') for _ in range(80)] + T2[90:]
    got = tsc.perform_alignment_batch([(T2, O2), (list('dominus'), list('dns')), (T, O)])
    assert tuple(got[0]) == tuple(py_port.perform_alignment(T2, O2))
    assert (''.join(got[1][0]), ''.join(got[1][1])) == ('domi_nus', '____dns_')
    assert tuple(got[2]) == tuple(py_port.perform_alignment(T, O))


def test_two_char_elements_demo_shape(tsc, oracle):
    """Elements need not be characters (textSeqCompare.py:185-186)."""
    rng = random.Random(4)
    T = [rng.choice(['Lo', 're', 'm ', 'ip', 'su']) for _ in range(90)]
    O = [rng.choice(['Lo', 're', 'm ', 'ip', 'xx']) for _ in range(120)]
    assert tuple(tsc.perform_alignment(T, O, [10, -5, -7, -7])) == tuple(oracle.perform_alignment(T, O, [10, -5, -7, -7]))


def test_unicode_symbols(tsc, oracle):
    T = list('dūs dominus ā ē alleluia ō') * 4
    O = list('dns dominvs a e allelvia ō') * 4
    assert tuple(tsc.perform_alignment(T, O)) == tuple(oracle.perform_alignment(T, O))


def test_inputs_not_mutated_and_outputs_fresh(tsc):
    T, O = list('gloria'), list('glorla')
    t0, o0 = list(T), list(O)
    a, b = tsc.perform_alignment(T, O)
    assert T == t0 and O == o0 and len(a) == len(b)
    a.append('x')
    assert tsc.perform_alignment(T, O)[0] != a


def test_verbose_prints_reference_format(tsc):
    buf = io.StringIO()
    with redirect_stdout(buf):
        tra, ocr = tsc.perform_alignment(list('ab'), list('ba'), verbose=True)
    lines = buf.getvalue().splitlines()
    assert (tra, ocr) == (list('ab_'), list('_ba'))
    assert lines == ['a _  ', 'b b O', '_ a  ']            # textSeqCompare.py:172-175


def test_invalid_scoring_system_raises_before_native(tsc):
    with pytest.raises(ValueError) as ei:
        tsc.perform_alignment(list('a'), list('b'), scoring_system=[1, 2, 3])
    assert str(ei.value) == 'scoring_system [1, 2, 3] invalid'


def test_score_range_guard(tsc):
    with pytest.raises(OverflowError):
        tsc.perform_alignment(list('ab'), list('ab'), scoring_system=[2 ** 24, -4, -7, -7, -3, 0])


def test_c2_sample_pages_bit_exact(tsc, oracle):
    """BASELINE config 2 shape: 48 of the 10k seeded page pairs vs the C oracle."""
    pairs = [synth.c2_pair(k) for k in range(48)]
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_c3_lines_bit_exact(tsc, oracle):
    pairs = [synth.c3_pair(k) for k in range(10000)]
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_c4_long_insertions_bit_exact(tsc, oracle):
    pairs = [synth.c4_pair(k) for k in range(24)]
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_determinism_and_order_independence(tsc):
    pairs = [synth.c3_pair(k) for k in range(200)] + [synth.c2_pair(k) for k in range(4)]
    a = tsc.align_packed(*_pack(pairs), DEFAULT)
    b = tsc.align_packed(*_pack(pairs), DEFAULT)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    rev = pairs[::-1]
    c = tsc.align_packed(*_pack(rev), DEFAULT)
    for k in range(len(pairs)):
        j = len(pairs) - 1 - k
        assert np.array_equal(a[0][a[1][k]:a[1][k] + a[2][k]], c[0][c[1][j]:c[1][j] + c[2][j]])


def test_op_string_invariants_at_scale(tsc):
    """Size-independent properties on a larger batch: every op string consumes exactly n
    transcript and m OCR symbols, and the recomputed path score equals the reported corner
    score of the state the traceback started in... (sum of per-column scores)."""
    pairs = [synth.c2_pair(1000 + k) for k in range(64)]
    buf, t_off, n, o_off, m = _pack(pairs)
    ops, ops_off, ops_len, scores = tsc.align_packed(buf, t_off, n, o_off, m, DEFAULT)
    for k in range(len(pairs)):
        o = ops[ops_off[k]:ops_off[k] + ops_len[k]]
        assert int((o != 2).sum()) == n[k] and int((o != 1).sum()) == m[k]
        assert max(n[k], m[k]) <= o.size <= n[k] + m[k]


# ---- whole-manuscript pairs: chained-pass path (BASELINE config 5) ------------------------------

@pytest.fixture()
def long_ctx(tsc):
    """A private context whose long-pair threshold is 1 cell, so every pair takes the
    chained-pass path regardless of size."""
    from text_alignment_b200 import _native
    ctx = _native.Context(0)
    ctx.set_long_threshold(1)
    yield ctx
    ctx.close()


def _check_ctx_vs_oracle(ctx, oracle, pairs, params=DEFAULT):
    buf, t_off, n, o_off, m = _pack(pairs)
    ops, ops_off, ops_len, scores = ctx.align_batch(buf, t_off, n, o_off, m, ctx.make_scoring(*params))
    sc, _ = oracle.make_scoring(list(params[:6]), boundary_gap=params[6])
    r_ops, r_off, r_len, r_end = oracle.align_batch_codes(buf, t_off, n, o_off, m, sc, threads=8)
    assert ops_len.tolist() == r_len.tolist()
    for k in range(len(pairs)):
        assert np.array_equal(ops[ops_off[k]:ops_off[k] + ops_len[k]], r_ops[r_off[k]:r_off[k] + r_len[k]]), \
            'pair %d (n=%d, m=%d)' % (k, n[k], m[k])
        got = tuple(None if v == -1073741824 else int(v) for v in scores[k].tolist())
        assert got == _end(r_end[k].tolist())


def test_long_path_small_and_ragged(long_ctx, oracle):
    pairs = [synth.make_pair(600 + k, n, m, 2, 30) for k, (n, m) in enumerate(
        [(1, 1), (1, 700), (700, 1), (31, 513), (33, 511), (64, 1024), (100, 1025), (300, 1100),
         (1500, 520), (40, 3000), (257, 2049)])]
    _check_ctx_vs_oracle(long_ctx, oracle, pairs)


def test_long_path_random_parameters(long_ctx, oracle):
    rng = random.Random(31)
    pairs = [(''.join(rng.choice('abc') for _ in range(rng.randint(1, 900))),
              ''.join(rng.choice('abc') for _ in range(rng.randint(1, 2600)))) for _ in range(6)]
    for params in [(5, -4, -2, -7, 0, -5, -1), (1, -1, -1, -1, -1, -1, -3), (7, 2, -4, -9, -1, 1, 0)]:
        _check_ctx_vs_oracle(long_ctx, oracle, pairs, params)


@pytest.mark.parametrize('rows', [1, 7, 32, 100, 257])
def test_long_path_row_bands(long_ctx, oracle, rows):
    """Row-banded chained-pass path (checkpoint at every band edge, bottom-up recompute with the
    traceback carried across bands): identical op strings and corner scores for every band
    height, including bands of one row and a last band shorter than the others."""
    long_ctx.set_long_band_rows(rows)
    sizes = [(1, 1), (2, 700), (700, 1), (33, 511), (100, 1025), (300, 1100), (801, 520), (257, 2049)]
    if rows < 32:
        sizes = sizes[:6]
    pairs = [synth.make_pair(900 + k, n, m, 2, 30) for k, (n, m) in enumerate(sizes)]
    _check_ctx_vs_oracle(long_ctx, oracle, pairs)
    rng = random.Random(rows)
    pairs = [(''.join(rng.choice('ab') for _ in range(rng.randint(1, 500))),
              ''.join(rng.choice('ab') for _ in range(rng.randint(1, 1500)))) for _ in range(4)]
    for params in [(5, -4, -2, -7, 0, -5, -1), (1, -1, -1, -1, -1, -1, -3), (7, 2, -4, 3, -1, 1, 0)]:
        _check_ctx_vs_oracle(long_ctx, oracle, pairs, params)


def test_long_path_row_bands_tabulated_scorer(long_ctx, oracle):
    import scorers
    long_ctx.set_long_band_rows(50)
    rng = random.Random(77)
    pairs = [(''.join(rng.choice('abcdef') for _ in range(420)), ''.join(rng.choice('abcdef') for _ in range(1300)))]
    buf, t_off, n, o_off, m = _pack(pairs)
    table = np.array([[scorers.SCORERS['vowel_aware'](chr(a), chr(b)) if a < 128 and b < 128 else -3
                       for b in range(128)] for a in range(128)], dtype=np.int32)
    sc = long_ctx.make_scoring(0, 0, -3, -4, -1, -2, -1, subst=table)
    ops, ops_off, ops_len, scores = long_ctx.align_batch(buf, t_off, n, o_off, m, sc)
    long_ctx.set_long_band_rows(0)
    ops1, _, ops_len1, scores1 = long_ctx.align_batch(buf, t_off, n, o_off, m, sc)
    assert ops_len.tolist() == ops_len1.tolist() and np.array_equal(ops[:ops_len[0]], ops1[:ops_len1[0]])
    assert scores.tolist() == scores1.tolist()


def test_pair_larger_than_the_arena_is_banded(oracle):
    """A pair whose pointer block (1 B/cell) exceeds the arena limit is aligned in row bands
    instead of being refused -- on the chained-pass path and for an oversized page of an ordinary
    batch alike."""
    from text_alignment_b200 import _native
    ctx = _native.Context(0)
    try:
        ctx.set_arena_limit(8 << 20)
        pairs = [synth.make_pair(5100, 9000, 8000, 5, 200)]            # 72 MB of pointers, 2^26 cells
        _check_ctx_vs_oracle(ctx, oracle, pairs)
        pairs = [synth.c2_pair(k) for k in range(6)] + [synth.make_pair(5101, 5000, 4000, 5, 200)]
        _check_ctx_vs_oracle(ctx, oracle, pairs)                       # 20 MB page among 2.7 MB pages
    finally:
        ctx.close()


def test_long_and_batched_pairs_mixed(tsc, oracle):
    """One batch holding ordinary pages and a pair above the default threshold (2^26 cells)."""
    pairs = [synth.c2_pair(7), synth.make_pair(5002, 9000, 8000, 5, 200), synth.c3_pair(3), synth.c2_pair(8)]
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_c5_whole_manuscript_pair(tsc, oracle):
    """BASELINE config 5 at full size: n=80 000 x m=100 000 (8e9 cells), bit-exact op string and
    corner scores against the C oracle (the Python reference would need 384 GB)."""
    t, o = synth.c5_pair()
    _check_packed_vs_oracle(tsc, oracle, [(t, o)], threads=1)


# ---- short pairs: four-pairs-per-warp line kernel (BASELINE config 3) ----------------------------

def test_line_kernel_boundaries_and_quads(tsc, oracle):
    """Strip-width classes (m = 32/64/96/128 edges), the m = 128/129 and n = 4096/4097 hand-over
    to the page kernel, quads padded with empty slots and mixed with cell-less pairs."""
    rng = random.Random(8)
    sizes = [(1, 1), (1, 128), (128, 1), (7, 32), (8, 33), (9, 64), (40, 65), (41, 96), (120, 97),
             (119, 128), (118, 129), (4096, 5), (4097, 5), (300, 100), (2, 2), (0, 7), (7, 0), (0, 0)]
    for count in (1, 2, 3, 5):            # incomplete quads
        pairs = [synth.make_pair(40 + k, n, m, 1, 6) for k, (n, m) in enumerate(sizes[:count])]
        _check_packed_vs_oracle(tsc, oracle, pairs)
    pairs = [synth.make_pair(70 + k, n, m, 1, 6) if n and m else ('a' * n, 'b' * m) for k, (n, m) in enumerate(sizes)]
    rng.shuffle(pairs)
    _check_packed_vs_oracle(tsc, oracle, pairs)


def test_line_kernel_matches_page_kernel(tsc, oracle):
    from text_alignment_b200 import _native
    pairs = [synth.c3_pair(50000 + k) for k in range(1500)]
    buf, t_off, n, o_off, m = _pack(pairs)
    a = tsc.align_packed(buf, t_off, n, o_off, m, DEFAULT)
    ctx = _native.Context(0)
    try:
        ctx.set_line_kernel(False)
        b = ctx.align_batch(buf, t_off, n, o_off, m, ctx.make_scoring(*DEFAULT))
    finally:
        ctx.close()
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    total = int(a[1][-1] + n[-1] + m[-1])
    for k in range(len(pairs)):
        assert np.array_equal(a[0][a[1][k]:a[1][k] + a[2][k]], b[0][b[1][k]:b[1][k] + b[2][k]])
    assert total == int((n.astype(np.int64) + m).sum())


def test_line_kernel_general_gap_extend_and_callable(tsc, oracle):
    """The gap_extend_y != 0 and substitution-table instantiations of the line kernel."""
    import scorers
    pairs = [synth.c3_pair(60000 + k) for k in range(300)]
    _check_packed_vs_oracle(tsc, oracle, pairs, (7, -3, -4, -9, -1, -2, -1))
    _check_packed_vs_oracle(tsc, oracle, pairs, (5, -4, -2, -7, 0, -5, -3))
    t, o = synth.make_pair(61000, 100, 120, 2, 6)
    system = [scorers.SCORERS['confusable'], -4, -6, -1, -2]
    got = tsc.perform_alignment(list(t), list(o), scoring_system=system, return_scores=True)
    want = oracle.perform_alignment(list(t), list(o), system, full=True)
    assert (got[0], got[1]) == (want[0], want[1]) and tuple(got[2]) == _end(want[2]['end'])


def test_parameter_sweep_resident_sequences(tsc, oracle):
    """evaluate_text_alignment.py:181-188: the same pages under many scoring vectors; the
    sequences are uploaded once and every vector is one rescore + launch."""
    from itertools import product
    grid = list(product([5, 8, 11], [-4, -7, -10], [-2, -5, -7], [-2, -5, -7], [0, -3, -5], [0, -3, -5]))
    rng = random.Random(3)
    systems = [list(p) for p in rng.sample(grid, 20)] + [[10, -5, -7, -7]]
    pairs = [(list(t), list(o)) for t, o in (synth.make_pair(900 + k, 150 + 30 * k, 200 + 25 * k, 2, 20) for k in range(3))]
    pairs.append((list('dominus'), list('dns')))
    got = tsc.perform_alignment_sweep(pairs, systems, return_scores=True)
    assert len(got) == len(systems)
    for system, res in zip(systems, got):
        for (T, O), (tra, ocr, score) in zip(pairs, res):
            want = oracle.perform_alignment(T, O, system, full=True)
            assert (tra, ocr) == (want[0], want[1]), system
            assert tuple(score) == _end(want[2]['end'])


def test_c2_all_10k_pages_bit_exact(tsc, oracle):
    """The north-star gate: all 10 000 seeded page pairs of BASELINE config 2, op strings and
    corner scores bit-exact against the C oracle (which is itself pinned to the reference)."""
    import os
    import bench
    cores = len(os.sched_getaffinity(0))
    packed, _ = bench.make_workload('c2', 0, 10000, max(1, min(cores, 32)))
    buf, t_off, n, o_off, m = packed
    ops, ops_off, ops_len, scores = tsc.align_packed(buf, t_off, n, o_off, m, DEFAULT)
    sc, _ = oracle.make_scoring(list(DEFAULT[:6]), boundary_gap=DEFAULT[6])
    r_ops, r_off, r_len, r_end = oracle.align_batch_codes(buf, t_off, n, o_off, m, sc, threads=cores)
    assert np.array_equal(ops_len, r_len) and np.array_equal(ops_off, r_off)
    assert np.array_equal(scores.astype(np.float64), np.where(r_end <= -1e99, -1073741824, r_end))
    used = np.zeros(ops.size, dtype=bool)
    idx = np.concatenate([np.arange(o, o + l) for o, l in zip(ops_off.tolist(), ops_len.tolist())])
    used[idx] = True
    assert np.array_equal(ops[used], r_ops[used])


def test_c4_256_pages_bit_exact(tsc, oracle):
    pairs = [synth.c4_pair(100 + k) for k in range(256)]
    _check_packed_vs_oracle(tsc, oracle, pairs, threads=16)


# ---- C-ABI error behaviour (int status + last_error, no aborts) ------------------

def test_abi_argument_errors(tsc):
    import ctypes
    from text_alignment_b200 import _native
    ctx = _native.Context(0)
    try:
        lib, h = ctx._lib, ctx._h
        sc, _ = ctx.make_scoring(*DEFAULT)
        assert lib.tanw_batch_run(h) == 6                                     # TANW_E_STATE
        assert b'before tanw_batch_prepare' in lib.tanw_last_error(h)
        buf = np.frombuffer(b'abcabd', dtype=np.uint8)
        good = dict(t_off=np.array([0], np.int64), n=np.array([3], np.int32),
                    o_off=np.array([3], np.int64), m=np.array([3], np.int32))
        with pytest.raises(ValueError):                                        # offsets outside the buffer
            ctx.align_batch(buf, good['t_off'], good['n'], np.array([5], np.int64), good['m'], (sc, None))
        with pytest.raises(ValueError):                                        # negative length
            ctx.align_batch(buf, good['t_off'], np.array([-1], np.int32), good['o_off'], good['m'], (sc, None))
        with pytest.raises(ValueError):                                        # table arrays differ in length
            ctx.align_batch(buf, good['t_off'], np.array([3, 3], np.int32), good['o_off'], good['m'], (sc, None))
        tab = np.zeros((2, 2), np.int32)                                       # symbol codes >= K
        with pytest.raises(ValueError):
            ctx.align_batch(buf, good['t_off'], good['n'], good['o_off'], good['m'],
                            ctx.make_scoring(0, 0, -1, -1, -1, -1, -1, subst=tab))
        # op buffer too small: the library refuses instead of writing out of bounds
        ctx.prepare(buf, good['t_off'], good['n'], good['o_off'], good['m'], (sc, None))
        ctx.run()
        ops = np.zeros(2, np.uint8); off = np.zeros(1, np.int64); ln = np.zeros(1, np.int32)
        u8p, i64p, i32p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32))
        rc = lib.tanw_batch_fetch(h, ops.ctypes.data_as(u8p), off.ctypes.data_as(i64p), 2, ln.ctypes.data_as(i32p), None)
        assert rc == 1 and b'op buffer too small' in lib.tanw_last_error(h)
        # a non-canonical op layout is honoured
        ops = np.full(40, 9, np.uint8); off = np.array([17], np.int64)
        rc = lib.tanw_batch_fetch(h, ops.ctypes.data_as(u8p), off.ctypes.data_as(i64p), 40, ln.ctypes.data_as(i32p), None)
        assert rc == 0 and ln[0] == 3 and ops[17:20].tolist() == [0, 0, 0] and ops[16] == 9 and ops[20] == 9
        # and the context still works afterwards
        out = ctx.align_batch(buf, good['t_off'], good['n'], good['o_off'], good['m'], (sc, None))
        assert out[0][:3].tolist() == [0, 0, 0]
    finally:
        ctx.close()


def test_empty_batch(tsc):
    z64, z32 = np.zeros(0, np.int64), np.zeros(0, np.int32)
    ops, off, ln, sc = tsc.align_packed(np.zeros(0, np.uint8), z64, z32, z64, z32, DEFAULT)
    assert ln.size == 0 and sc.shape == (0, 3)
    assert tsc.perform_alignment_batch([]) == []


def test_plain_c_caller(tmp_path):
    """examples/abi_example.c: the C ABI is usable without Python (SURVEY App. B KATs)."""
    import os
    import shutil
    import subprocess
    if not shutil.which('gcc'):
        pytest.skip('no gcc')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'abi_example')
    libdir = os.path.join(root, 'text_alignment_b200')
    subprocess.check_call(['gcc', '-I' + os.path.join(root, 'include'), os.path.join(root, 'examples', 'abi_example.c'),
                           '-o', exe, '-L' + libdir, '-ltanw', '-Wl,-rpath,' + libdir])
    out = subprocess.check_output([exe], text=True).splitlines()
    assert out == ['domi_nus', '____dns_', '(M, X, Y)[n][m] = (2, -10, -7)',
                   'allel_____u__ia', 'a l l e l u y a', '(M, X, Y)[n][m] = (14, -3, 2)']


def test_small_arena_limit_reduces_occupancy_not_correctness(oracle):
    """The pointer arena is per resident warp; a tight limit must only reduce the number of
    resident warps (or refuse a pair that cannot fit at all), never change results."""
    from text_alignment_b200 import _native
    ctx = _native.Context(0)
    try:
        pairs = [synth.c2_pair(300 + k) for k in range(12)] + [synth.c3_pair(k) for k in range(40)]
        ctx.set_arena_limit(24 << 20)                     # room for ~2 CTAs of page slots
        _check_ctx_vs_oracle(ctx, oracle, pairs)
        ctx.set_arena_limit(1 << 20)                      # smaller than one page: row bands
        _check_ctx_vs_oracle(ctx, oracle, pairs[:14])
        ctx.set_arena_limit(16 << 10)                     # not even 32 rows of one page
        with pytest.raises(MemoryError):
            ctx.align_batch(*_pack(pairs[:2]), ctx.make_scoring(*DEFAULT))
        ctx.set_arena_limit(0)                            # back to the default
        _check_ctx_vs_oracle(ctx, oracle, pairs[:3])
    finally:
        ctx.close()


# ---- round 2: one-launch sweep, batch-wide tables, pipelined batches, shared contexts ------------

def test_parameter_sweep_is_one_launch(tsc, oracle):
    """evaluate_text_alignment.py:134-194: 729 scoring vectors x 3 pages = 2187 independent
    alignments; here all of them are ONE batch with per-pair scoring systems
    (tanw_align_batch_multi), including vectors that need the general recurrences."""
    from itertools import product
    grid = [list(p) for p in product([5, 8, 11], [-4, -7, -10], [-2, -5, -7], [-2, -5, -7], [0, -3, -5], [0, -3, -5])]
    pages = [(list(t), list(o)) for t, o in (synth.make_pair(700 + k, 260 + 40 * k, 330 + 50 * k, 3, 25) for k in range(3))]
    got = tsc.perform_alignment_sweep(pages, grid, return_scores=True)
    tm = tsc.get_context(0).timing()
    assert tm['kernel_launches'] == 1 and tm['cells'] == 729 * sum(len(t) * len(o) for t, o in pages)
    rng = random.Random(11)
    for k in rng.sample(range(729), 40) + [0, 728]:
        for (T, O), (tra, ocr, score) in zip(pages, got[k]):
            want = oracle.perform_alignment(T, O, grid[k], full=True)
            assert (tra, ocr) == (want[0], want[1]), grid[k]
            assert tuple(score) == _end(want[2]['end'])
    # mixed variants in one batch: a positive gap open forces the general recurrences for all
    mixed = [[8, -4, -7, -7, -3, 0], [7, 2, 3, -4, -1, 1], [5, -4, -2, -7, 0, -5], [10, -5, -7, -7]]
    got = tsc.perform_alignment_sweep(pages + [(list('dominus'), list('dns')), ([], list('ab'))], mixed, return_scores=True)
    for system, res in zip(mixed, got):
        for (T, O), (tra, ocr, score) in zip(pages + [(list('dominus'), list('dns')), ([], list('ab'))], res):
            want = oracle.perform_alignment(T, O, system, full=True)
            assert (tra, ocr) == (want[0], want[1]) and tuple(score) == _end(want[2]['end']), system


def test_callable_scorer_batch_shares_one_table(tsc, oracle):
    """A callable scorer (textSeqCompare.py:27-29) over a batch: the batch is interned once, one
    K x K table serves every pair and the whole batch is one launch per kernel family."""
    def fn(a, b):
        return 6 if a == b else (-1 if (a in 'aeiou') == (b in 'aeiou') else -5)
    pairs = [(list(t), list(o)) for t, o in
             [synth.c3_pair(k) for k in range(150)] + [synth.make_pair(50 + k, 300, 420, 3, 30) for k in range(6)]]
    pairs += [([], list('ab')), (list('dominus'), list('dns'))]
    for system in ([fn, -7, -6, -3, -1], [fn, -7, -7, -3, 0], [fn, 2, -6, -3, -1]):
        got = tsc.perform_alignment_batch(pairs, system, return_scores=True)
        assert tsc.get_context(0).timing()['kernel_launches'] <= 2          # line kernel + page kernel
        for (T, O), (tra, ocr, score) in zip(pairs, got):
            want = oracle.perform_alignment(T, O, system, full=True)
            assert (tra, ocr) == (want[0], want[1]), system[1:]
            assert tuple(score) == _end(want[2]['end'])


def test_pipelined_batch_equals_three_phase(tsc, oracle):
    """tanw_align_batch cuts a copy-heavy batch into chunks (uploads, kernels and downloads
    overlap); the three-phase form runs the same batch as one chunk.  Same bytes either way."""
    from text_alignment_b200 import _native
    pairs = [synth.c3_pair(k) for k in range(70000)]
    pairs[12345] = ('', 'abc')
    pairs[40000] = synth.make_pair(5, 150, 210, 5, 40)         # a small page among the lines
    buf, t_off, n, o_off, m = _pack(pairs)
    ctx = _native.Context(0)
    try:
        sc = ctx.make_scoring(*DEFAULT)
        one = ctx.align_batch(buf, t_off, n, o_off, m, sc)
        assert ctx.timing()['chunks'] > 1
        ctx.prepare(buf, t_off, n, o_off, m, sc)
        ctx.run()
        three = ctx.fetch()
        assert ctx.timing()['chunks'] == 1
        assert np.array_equal(one[2], three[2]) and np.array_equal(one[3], three[3])
        valid = np.zeros(one[0].size + 1, dtype=np.int32)
        np.add.at(valid, one[1], 1)
        np.add.at(valid, one[1] + one[2], -1)
        mask = np.cumsum(valid[:-1]) > 0
        assert np.array_equal(one[0][mask], three[0][:mask.size][mask])
    finally:
        ctx.close()
    sample = list(range(0, 70000, 997)) + [12345, 40000]
    osc, _ = oracle.make_scoring(list(DEFAULT[:6]), boundary_gap=DEFAULT[6])
    sub = _pack([pairs[k] for k in sample])
    r_ops, r_off, r_len, r_end = oracle.align_batch_codes(*sub, osc, threads=8)
    for j, k in enumerate(sample):
        assert np.array_equal(one[0][one[1][k]:one[1][k] + one[2][k]], r_ops[r_off[j]:r_off[j] + r_len[j]]), k
        assert tuple(None if v == -1073741824 else int(v) for v in one[3][k].tolist()) == _end(r_end[j].tolist())


def test_two_threads_share_one_context(tsc, oracle):
    """The reference function is pure and re-entrant; here two Python threads calling the
    drop-in at once share the device's context and are serialised by its lock."""
    import threading
    jobs = [(list(t), list(o)) for t, o in (synth.make_pair(300 + k, 120 + 7 * k, 150 + 5 * k, 2, 12) for k in range(24))]
    want = [oracle.perform_alignment(T, O, None) for T, O in jobs]
    results = {}
    errors = []

    def work(tid):
        try:
            for rep in range(6):
                for k in range(tid, len(jobs), 2):
                    results[(tid, rep, k)] = tsc.perform_alignment(jobs[k][0], jobs[k][1])
        except BaseException as e:       # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors
    assert len(results) == 6 * len(jobs)
    for (tid, rep, k), got in results.items():
        assert (got[0], got[1]) == (want[k][0], want[k][1])


def test_wide_symbol_pair_beyond_the_arena_is_banded(oracle):
    """16-bit symbol codes on the chained-stripe path: a pair whose pointers exceed one warp's
    share of the arena is cut into row bands like any other (round 1 refused it)."""
    from text_alignment_b200 import _native
    t, o = synth.make_pair(9, 700, 900, 3, 30)
    buf, t_off, n, o_off, m = _pack([(t, o), ('abc', 'abd')])
    ctx = _native.Context(0)
    try:
        narrow = ctx.align_batch(buf, t_off, n, o_off, m, ctx.make_scoring(*DEFAULT))
        ctx.set_arena_limit(512 << 10)
        for params in (DEFAULT, (7, 2, -4, 3, -1, 1, 0)):
            narrow = ctx.align_batch(buf, t_off, n, o_off, m, ctx.make_scoring(*params))
            wide = ctx.align_batch(buf.astype(np.uint16), t_off, n, o_off, m, ctx.make_scoring(*params))
            assert np.array_equal(narrow[2], wide[2]) and np.array_equal(narrow[3], wide[3])
            for k in range(2):
                assert np.array_equal(narrow[0][narrow[1][k]:narrow[1][k] + narrow[2][k]],
                                      wide[0][wide[1][k]:wide[1][k] + wide[2][k]])
    finally:
        ctx.close()
    _ = oracle


def _same_results(a, b, count):
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    for k in range(count):
        assert np.array_equal(a[0][a[1][k]:a[1][k] + a[2][k]], b[0][b[1][k]:b[1][k] + b[2][k]]), k


def test_line16_kernel_matches_int32_line_kernel_and_oracle(oracle):
    """The .u16x2 line kernel (two pairs per register, eight per warp) against the int32 line
    kernel on the same batch and against the oracle: strip-width class edges, octets padded with
    empty slots, cell-less pairs, heights either side of the 16-bit range limit (taller pairs
    stay on the int32 kernel), gap_extend_y != 0 (variant 1), and scoring systems that are not
    eligible at all (general recurrences, mismatch > match)."""
    from text_alignment_b200 import _native
    sizes = [(1, 1), (1, 128), (128, 1), (7, 32), (8, 33), (9, 64), (40, 65), (41, 96), (120, 97), (119, 128),
             (334, 90), (335, 90), (333, 128), (400, 60), (2, 2), (0, 7), (7, 0), (0, 0), (64, 64), (63, 64), (65, 64)]
    pairs = [synth.make_pair(90 + k, n, m, 1, 6) if n and m else ('a' * n, 'b' * m) for k, (n, m) in enumerate(sizes)]
    pairs += [synth.c3_pair(70000 + k) for k in range(3000)]
    random.Random(4).shuffle(pairs)
    buf, t_off, n, o_off, m = _pack(pairs)
    ctx = _native.Context(0)
    try:
        for params in (DEFAULT, (7, -3, -4, -9, -1, -2, -1), (5, -4, -2, -7, 0, -5, -3), (11, -10, -7, -7, -5, -5, -1),
                       (7, 2, 3, -4, -1, 1, 0), (-4, 8, -7, -7, -3, 0, -1)):
            sc = ctx.make_scoring(*params)
            ctx.set_line_kernel(1)
            a = ctx.align_batch(buf, t_off, n, o_off, m, sc)
            launches = ctx.timing()['kernel_launches']
            ctx.set_line_kernel(2)
            b = ctx.align_batch(buf, t_off, n, o_off, m, sc)
            _same_results(a, b, len(pairs))
            eligible = params[2] <= 0 and params[3] <= 0 and params[0] >= params[1]
            # 16-bit fill + traceback kernels (if eligible) + int32 line kernel (the tall pairs) 
            assert (launches >= 2) if eligible else (launches <= 2), (params, launches)
            osc, _ = oracle.make_scoring(list(params[:6]), boundary_gap=params[6])
            r_ops, r_off, r_len, r_end = oracle.align_batch_codes(buf, t_off, n, o_off, m, osc, threads=8)
            assert a[2].tolist() == r_len.tolist()
            for k in range(len(pairs)):
                assert np.array_equal(a[0][a[1][k]:a[1][k] + a[2][k]], r_ops[r_off[k]:r_off[k] + r_len[k]]), (params, k)
                got = tuple(None if v == -1073741824 else int(v) for v in a[3][k].tolist())
                assert got == _end(r_end[k].tolist()), (params, k, n[k], m[k])
        # incomplete octets
        for count in (1, 2, 3, 5, 9, 17):
            sub = _pack(pairs[:count])
            ctx.set_line_kernel(1)
            a = ctx.align_batch(*sub, ctx.make_scoring(*DEFAULT))
            ctx.set_line_kernel(0)
            b = ctx.align_batch(*sub, ctx.make_scoring(*DEFAULT))
            _same_results(a, b, count)
    finally:
        ctx.close()


def test_line16_rescore_needs_an_eligible_system():
    from text_alignment_b200 import _native
    pairs = [synth.c3_pair(k) for k in range(64)]
    ctx = _native.Context(0)
    try:
        ctx.prepare(*_pack(pairs), ctx.make_scoring(*DEFAULT))
        ctx.run()
        a = ctx.fetch()
        ctx.rescore(ctx.make_scoring(5, -4, -2, -7, 0, -5, -1))        # still eligible
        ctx.run()
        ctx.fetch()
        with pytest.raises(_native.NativeError):                     # a positive gap open: prepare again
            ctx.rescore(ctx.make_scoring(7, 2, 3, -4, -1, 1, 0))
        ctx.rescore(ctx.make_scoring(*DEFAULT))
        ctx.run()
        _same_results(a, ctx.fetch(), len(pairs))
    finally:
        ctx.close()


def test_packed_ops_are_the_same_alignments(oracle):
    """tanw_set_packed_ops: four ops per byte on the way back (a quarter of the PCIe traffic);
    unpacked, they are the bytes the default mode returns -- for lines (both line kernels), pages,
    a chained-stripe pair, cell-less pairs, through the chunked pipeline and the three-phase form."""
    from text_alignment_b200 import _native
    pairs = [synth.c3_pair(k) for k in range(45000)] + [synth.c2_pair(k) for k in range(6)] + \
            [('', 'abc'), ('abc', ''), ('', ''), synth.make_pair(3, 400, 90, 2, 6)]
    buf, t_off, n, o_off, m = _pack(pairs)
    ctx = _native.Context(0)
    try:
        ctx.set_long_threshold(2500000)                  # the largest page takes the chained-stripe path
        sc = ctx.make_scoring(*DEFAULT)
        plain = ctx.align_batch(buf, t_off, n, o_off, m, sc)
        ctx.set_packed_ops(True)
        packed = ctx.align_batch(buf, t_off, n, o_off, m, sc)             # one chunk: a chained-stripe pair is in it
        assert packed[0].size == int((n.astype(np.int64) + m).sum()) // 4 + len(pairs) + 1
        ctx.prepare(buf, t_off, n, o_off, m, sc)
        ctx.run()
        three = ctx.fetch()
        ctx.set_long_threshold(1 << 26)
        piped = ctx.align_batch(buf, t_off, n, o_off, m, sc)              # without it: the chunked pipeline
        assert ctx.timing()['chunks'] > 1
        ctx.set_packed_ops(False)
        for got in (packed, three, piped):
            assert np.array_equal(got[2], plain[2]) and np.array_equal(got[3], plain[3])
            ops, off = ctx.unpack_ops(got[0], n, m, got[2])
            assert np.array_equal(off, plain[1])
            for k in list(range(0, len(pairs), 211)) + list(range(len(pairs) - 10, len(pairs))):
                assert np.array_equal(ops[off[k]:off[k] + got[2][k]], plain[0][plain[1][k]:plain[1][k] + plain[2][k]]), k
    finally:
        ctx.close()
    _ = oracle


def test_callable_scorer_profile_and_lookup_paths(tsc, oracle):
    """A tabulated scorer runs the page kernel on a per-lane query profile (signed bytes) when the
    alphabet has at most 32 symbols and |score| <= 127, and on table lookups otherwise: both
    against the oracle, with remainder passes of every strip width."""
    def small(a, b):
        return 7 if a == b else (-1 if (a in 'aeiou') == (b in 'aeiou') else -6)

    def big(a, b):
        return 200 if a == b else -150
    rng = random.Random(21)
    pages = [synth.make_pair(800 + k, 200 + 37 * k, 150 + 61 * k, 3, 30) for k in range(8)]         # m = 150 .. 577
    wide_alpha = [chr(0x61 + k) for k in range(26)] + [chr(0x3b1 + k) for k in range(20)]           # 46 symbols
    wide = [([rng.choice(wide_alpha) for _ in range(180)], [rng.choice(wide_alpha) for _ in range(260)]) for _ in range(3)]
    for fn, pairs in ((small, [(list(t), list(o)) for t, o in pages]), (big, [(list(t), list(o)) for t, o in pages[:4]]),
                      (small, wide)):
        for gaps in ([-7, -6, -3, -1], [-7, -7, -3, 0], [2, -6, -3, -1]):
            system = [fn] + gaps
            got = tsc.perform_alignment_batch(pairs, system, return_scores=True)
            for (T, O), (tra, ocr, score) in zip(pairs, got):
                want = oracle.perform_alignment(T, O, system, full=True)
                assert (tra, ocr) == (want[0], want[1]), (fn.__name__, gaps)
                assert tuple(score) == _end(want[2]['end'])
