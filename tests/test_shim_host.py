"""Host-side logic of the drop-in module that needs no GPU: scoring-system parsing
(textSeqCompare.py:24-42), element interning, op decoding, sharding."""
import numpy as np
import pytest

from text_alignment_b200 import textSeqCompare as tsc


def test_module_constants_match_reference():
    # textSeqCompare.py:6-10
    assert (tsc.default_match, tsc.default_mismatch, tsc.gap_open, tsc.gap_extend) == (10, -5, -10, -1)
    assert tsc.default_sys == [8, -4, -7, -7, -3, 0]


def test_parse_scoring_forms():
    assert tsc.parse_scoring_system(None) == (None, 8, -4, -7, -7, -3, 0)
    assert tsc.parse_scoring_system([10, -5, -7, -7]) == (None, 10, -5, -7, -7, -7, -7)
    assert tsc.parse_scoring_system(np.array([5, -4, -2, -7, 0, -5])) == (None, 5, -4, -2, -7, 0, -5)
    f = lambda a, b: 1
    assert tsc.parse_scoring_system([f, -1, -2, -3, -4]) == (f, 0, 0, -1, -2, -3, -4)


def test_default_sys_read_at_call_time(monkeypatch):
    monkeypatch.setattr(tsc, 'default_sys', [1, -1, -2, -2])
    assert tsc.parse_scoring_system(None) == (None, 1, -1, -2, -2, -2, -2)


@pytest.mark.parametrize('bad', [[1, 2, 3], [], [1, 2, 3, 4, 5], [1] * 7])
def test_invalid_scoring_system_message(bad):
    with pytest.raises(ValueError) as ei:
        tsc.parse_scoring_system(bad)
    assert str(ei.value) == 'scoring_system {} invalid'.format(bad)      # textSeqCompare.py:42


def test_non_integral_scoring_rejected():
    with pytest.raises(TypeError):
        tsc.parse_scoring_system([8.5, -4, -7, -7, -3, 0])
    assert tsc.parse_scoring_system([8.0, -4, -7, -7, -3, 0])[1] == 8


def test_inputs_must_be_lists_like_the_reference():
    # the reference does `transcript + [' ']` (:21), which raises for str and tuple
    with pytest.raises(TypeError):
        tsc.perform_alignment_batch([('abc', list('abc'))])
    with pytest.raises(TypeError):
        tsc.perform_alignment_batch([(list('abc'), tuple('abc'))])


def _codes(enc, k):
    """(transcript codes, OCR codes) of pair k of an encoded batch."""
    t = enc.symbols[enc.t_off[k]:enc.t_off[k] + enc.n[k]]
    o = enc.symbols[enc.o_off[k]:enc.o_off[k] + enc.m[k]]
    return t, o


def test_encode_single_chars_uses_code_points():
    enc = tsc._encode_batch([(list('gloria'), list('glorla'))], tabulated=False)
    assert enc.alphabet is None and _codes(enc, 0)[0].tolist() == [ord(c) for c in 'gloria']
    enc = tsc._encode_batch([(list('dūs'), list('dns'))], tabulated=False)       # non-latin1 symbol
    assert enc.alphabet is not None
    t, o = _codes(enc, 0)
    assert [enc.alphabet[c] for c in t] == list('dūs')
    assert [enc.alphabet[c] for c in o] == list('dns')


def test_encode_is_batch_wide_and_helper_agrees_with_numpy(monkeypatch):
    """One interning for the whole batch (one launch, one table), and the CPython helper
    (csrc/tanw_pylist.c) gives what the pure numpy route gives."""
    pairs = [(list('gloria'), list('glorla')), ([], list('x')), (list('dūs'), []), (list('in excelsis'), list('ln exce1sis'))]
    with_helper = tsc._encode_batch(pairs, tabulated=True)
    monkeypatch.setattr(tsc._native, 'pylist', lambda: None)
    without = tsc._encode_batch(pairs, tabulated=True)
    for enc in (with_helper, without):
        assert enc.symbols.dtype == np.uint8 and enc.n.tolist() == [6, 0, 3, 11] and enc.m.tolist() == [6, 1, 0, 11]
        for k, (t, o) in enumerate(pairs):
            ct, co = _codes(enc, k)
            assert [enc.alphabet[c] for c in ct] == t and [enc.alphabet[c] for c in co] == o
    assert np.array_equal(with_helper.symbols, without.symbols) and with_helper.alphabet == without.alphabet
    ops = np.array([0, 0, 2, 1, 0], dtype=np.uint8)
    assert tsc._decode(list('abcd'), list('abxd'), ops) == (list('ab_cd'), list('abx_d'))


def test_encode_general_elements():
    T = ['Lo', 're', 'm ', 7, (1, 2), 'Lo']
    O = ['re', 7.0, (1, 2), [1], [1]]
    enc = tsc._encode_batch([(T, O)], tabulated=False)
    t, o = _codes(enc, 0)
    assert t[0] == t[5]
    assert t[1] == o[0]
    assert t[3] == o[1]          # 7 == 7.0
    assert t[4] == o[2]
    assert o[3] == o[4]          # unhashable but equal
    assert enc.reflexive
    enc = tsc._encode_batch([([float('nan')], [1.0])], tabulated=False)
    assert not enc.reflexive


def test_wide_alphabets_get_16_bit_codes():
    """More than 256 distinct elements in one launch: uint16 codes (page kernel on the device);
    a callable scorer is tabulated, which bounds the alphabet at 2048; beyond that the batch is
    encoded pair by pair (None)."""
    t = [chr(0x400 + k) for k in range(300)]
    enc = tsc._encode_batch([(t, list('ab') + t[:5])], tabulated=False)
    assert enc.symbols.dtype == np.uint16
    assert len(enc.alphabet) == 302
    ct, co = _codes(enc, 0)
    assert [enc.alphabet[c] for c in ct.tolist()] == t
    assert [enc.alphabet[c] for c in co.tolist()] == list('ab') + t[:5]
    # non-string elements take the dictionary path
    enc = tsc._encode_batch([([(k, k) for k in range(400)], [(3, 3), (500, 1)])], tabulated=False)
    assert enc.symbols.dtype == np.uint16 and enc.alphabet[_codes(enc, 0)[1][1]] == (500, 1)
    # narrow pairs stay 8 bits wide
    assert tsc._encode_batch([(list('abc'), [chr(0x400)])], tabulated=False).symbols.dtype == np.uint8
    assert tsc._encode_batch([([chr(0x400 + k) for k in range(3000)], list('ab'))], tabulated=True) is None
    assert tsc._encode_batch([(list(range(70000)), [1])], tabulated=False) is None


def test_tabulate_only_calls_needed_pairs():
    seen = set()

    def f(a, b):
        seen.add((a, b))
        return 3 if a == b else -2
    enc = tsc._encode_batch([(list('abca'), list('xa')), (list('q'), list('ab'))], tabulated=True)
    tab = tsc._tabulate(enc, f, 0, 0)
    assert seen == {(a, b) for a in 'abc' for b in 'xa'} | {('q', 'a'), ('q', 'b')}
    ia, ix = enc.alphabet.index('a'), enc.alphabet.index('x')
    assert tab[ia, ia] == 3 and tab[ia, ix] == -2


def test_decode_ops():
    T, O = list('dominus'), list('dns')
    tra, oc = 'domi_nus', '____dns_'
    ops = np.array([1 if b == '_' else (2 if a == '_' else 0) for a, b in zip(tra, oc)], dtype=np.uint8)
    assert tsc._decode(T, O, ops) == (list(tra), list(oc))
    assert ''.join(tsc._align_record(T, O, ops)) == '     O~ '


def test_decode_keeps_caller_objects(monkeypatch):
    a, b = ('x', 1), ('x', 1)
    T, O = [a, 'q'], [b]
    for helper in (True, False):
        if not helper:
            monkeypatch.setattr(tsc._native, 'pylist', lambda: None)
        tra, oc = tsc._decode(T, O, np.array([0, 1], dtype=np.uint8))
        assert tra[0] is a and oc[0] is b and oc[1] == '_'
    with pytest.raises((ValueError, StopIteration)):
        tsc._decode(T, O, np.array([0, 0, 0], dtype=np.uint8))


def test_split_by_cells_is_balanced_partition():
    rng = np.random.default_rng(3)
    n = rng.integers(1000, 1600, size=999).astype(np.int32)
    m = (n * 1.25).astype(np.int32)
    b = tsc.split_by_cells(n, m, 8)
    assert b[0] == 0 and b[-1] == 999 and np.all(np.diff(b) >= 0)
    cells = n.astype(np.int64) * m
    loads = [cells[b[k]:b[k + 1]].sum() for k in range(8)]
    assert max(loads) / (sum(loads) / 8) < 1.02
    assert tsc.split_by_cells(n[:3], m[:3], 8)[-1] == 3
    assert tsc.split_by_cells(n[:0], m[:0], 4).tolist() == [0, 0, 0, 0, 0]


def test_gather_shards_layout():
    n = np.array([2, 0, 3, 1], dtype=np.int32)
    m = np.array([1, 0, 2, 4], dtype=np.int32)
    # shard 0 = pairs 0..1, shard 1 = pairs 2..3; fake per-shard outputs in canonical layout
    s0 = (np.array([0, 1, 9], np.uint8), np.array([0, 3]), np.array([2, 0], np.int32), np.zeros((2, 3), np.int32))
    s1 = (np.arange(10, dtype=np.uint8), np.array([0, 5]), np.array([5, 4], np.int32), np.ones((2, 3), np.int32))
    ops, off, ln, sc = tsc.gather_shards([s0, s1], n, m, True, np.array([0, 2, 4]))
    assert off.tolist() == [0, 3, 3, 8]
    assert ops[:3].tolist() == [0, 1, 9] and ops[3:13].tolist() == list(range(10))
    assert ln.tolist() == [2, 0, 5, 4] and sc.shape == (4, 3)
