"""Pin the oracle (oracle/nw_oracle.c and oracle/py_port.py) against the reference.

Golden vectors were produced by the unmodified reference (tests/golden/make_golden.py); when
/root/reference is present (build container) the live reference is run as well."""
import hashlib
import random

import numpy as np
import pytest

from conftest import golden_elems, ops_string, resolve_system
from oracle import nw_oracle, py_port, ref_loader
from text_alignment_b200 import synth


def _end_ints(end):
    return [None if v <= -1e99 else int(v) for v in end]


def _check_c_oracle(rec):
    T, O = golden_elems(rec)
    tra, ocr, full = nw_oracle.perform_alignment(T, O, resolve_system(rec['system']), full=True)
    assert ''.join(map(str, full['ops'].tolist())) == rec['ops']
    assert _end_ints(full['end']) == rec['end']
    got = hashlib.sha256(np.ascontiguousarray(full['ptr']).tobytes()).hexdigest()
    assert got == rec['ptr_sha256']
    return tra, ocr


def test_c_oracle_kats(kats):
    for rec in kats['kats']:
        tra, ocr = _check_c_oracle(rec)
        assert ''.join(tra) == rec['tra'] and ''.join(ocr) == rec['ocr']
    tra, ocr = _check_c_oracle(kats['demo'])
    assert '|'.join(tra) == kats['demo']['tra'] and '|'.join(ocr) == kats['demo']['ocr']
    tra, ocr = _check_c_oracle(kats['demo_chars'])
    assert ''.join(tra) == kats['demo_chars']['tra']


def test_quirk_discriminators(kats):
    """a textbook Gotoh implementation gets these three wrong."""
    by = {(r['T'], r['O']): r for r in kats['kats'] if r['system'] is None}
    assert (by[('ca', 'aa')]['tra'], by[('ca', 'aa')]['ocr']) == ('ca_', '_aa')
    assert (by[('a', 'a')]['tra'], by[('a', 'a')]['ocr']) == ('a', 'a')
    assert (by[('a', 'c')]['tra'], by[('a', 'c')]['ocr']) == ('a', 'c')


def test_c_oracle_random_golden(random_pairs):
    for rec in random_pairs:
        _check_c_oracle(rec)


def test_py_port_golden(kats, random_pairs):
    recs = kats['kats'] + [kats['demo']] + random_pairs[:160] + random_pairs[-3:-1]
    for rec in recs:
        T, O = golden_elems(rec)
        tra, ocr, full = py_port.perform_alignment(T, O, resolve_system(rec['system']), full=True)
        assert ''.join(map(str, full['ops'])) == rec['ops']
        n, m = len(T), len(O)
        end = [full[k][n][m] for k in ('M', 'X', 'Y')]
        assert _end_ints(end) == rec['end']
        ptr = (full['PM'].astype(np.uint8) | (full['PX'].astype(np.uint8) << 2) |
               (full['PY'].astype(np.uint8) << 4))[1:, 1:]
        assert hashlib.sha256(np.ascontiguousarray(ptr).tobytes()).hexdigest() == rec['ptr_sha256']


def test_py_port_wide_alphabet_golden(wide_pairs):
    """Vectors of the unmodified reference on pairs with 326-409 distinct elements (the C oracle
    is limited to 256 codes; the Python port is the checker for the 16-bit device path)."""
    for rec in wide_pairs:
        T, O = rec['T'], rec['O']
        tra, ocr, full = py_port.perform_alignment(T, O, resolve_system(rec['system']), full=True)
        assert ''.join(map(str, full['ops'])) == rec['ops']
        end = [full[k][len(T)][len(O)] for k in ('M', 'X', 'Y')]
        assert _end_ints(end) == rec['end']
        ptr = (full['PM'].astype(np.uint8) | (full['PX'].astype(np.uint8) << 2) |
               (full['PY'].astype(np.uint8) << 4))[1:, 1:]
        assert hashlib.sha256(np.ascontiguousarray(ptr).tobytes()).hexdigest() == rec['ptr_sha256']


def test_c_oracle_appendix_c(appendix_c):
    for rec in appendix_c:
        t, o = synth.make_pair(rec['seed'], rec['n'], rec['m'], rec['run_lo'], rec['run_hi'])
        tra, ocr, full = nw_oracle.perform_alignment(list(t), list(o), None, full=True)
        assert hashlib.sha256((''.join(tra) + '\n' + ''.join(ocr)).encode()).hexdigest() == rec['align_sha256']
        assert hashlib.sha256(np.ascontiguousarray(full['ptr']).tobytes()).hexdigest() == rec['ptr_sha256']
        assert _end_ints(full['end']) == rec['end']
        assert len(full['ops']) == rec['L']
        assert tra.count('_') == rec['gaps_tra'] and ocr.count('_') == rec['gaps_ocr']


def test_batch_matches_single():
    rng = random.Random(5)
    pairs = [synth.make_pair(9000 + k, rng.randint(0, 90), rng.randint(0, 90), 1, 5) for k in range(24)]
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode(), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs]); m = np.array([len(o) for _, o in pairs])
    t_off = np.cumsum(np.concatenate([[0], (n + m)[:-1]])); o_off = t_off + n
    sc, _ = nw_oracle.make_scoring(None)
    ops, ops_off, ops_len, end3 = nw_oracle.align_batch_codes(buf, t_off, n, o_off, m, sc, threads=3)
    for k, (t, o) in enumerate(pairs):
        _, _, full = nw_oracle.perform_alignment(list(t), list(o), None, full=True)
        assert ops[ops_off[k]:ops_off[k] + ops_len[k]].tolist() == full['ops'].tolist()
        assert tuple(end3[k].tolist()) == full['end']


def test_boundary_gap_is_a_separate_parameter():
    """The boundary rows use the module constant gap_extend (textSeqCompare.py:9, :54-59),
    not the call's gap parameters: changing it changes results independently."""
    a = nw_oracle.perform_alignment(list('ca'), list('aa'), None, boundary_gap=-1)
    b = nw_oracle.perform_alignment(list('ca'), list('aa'), None, boundary_gap=-9)
    assert (''.join(a[0]), ''.join(a[1])) == ('ca_', '_aa')
    assert a != b


@pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not present')
def test_live_reference_differential():
    """Live differential run against the unmodified reference (build container only)."""
    tsc = ref_loader.load_textseqcompare()
    rng = random.Random(77)
    for k in range(300):
        alpha = rng.choice(['ab', 'acgt', 'aeiou dnm'])
        T = [rng.choice(alpha) for _ in range(rng.randint(0, 26))]
        O = [rng.choice(alpha) for _ in range(rng.randint(0, 26))]
        s = None if k % 2 else [rng.randint(0, 9), rng.randint(-9, 1)] + [rng.randint(-9, 1) for _ in range(4)]
        ref = tsc.perform_alignment(list(T), list(O), scoring_system=s)
        assert nw_oracle.perform_alignment(T, O, s) == ref
        if k % 10 == 0:
            assert tuple(py_port.perform_alignment(T, O, s)) == ref


@pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not present')
def test_live_reference_module_gap_extend():
    """gap_extend is read at call time (textSeqCompare.py:54-59)."""
    tsc = ref_loader.load_textseqcompare()
    saved = tsc.gap_extend
    try:
        for g in (-1, -4, 0, 2):
            tsc.gap_extend = g
            for T, O in [('ca', 'aa'), ('dominus', 'dns'), ('abcabc', 'abc')]:
                ref = tsc.perform_alignment(list(T), list(O))
                assert nw_oracle.perform_alignment(list(T), list(O), None, boundary_gap=g) == ref
    finally:
        tsc.gap_extend = saved


def test_invalid_scoring_system():
    with pytest.raises(ValueError):
        nw_oracle.make_scoring([1, 2, 3])
    with pytest.raises(ValueError):
        py_port.perform_alignment(list('a'), list('b'), [1, 2, 3])
