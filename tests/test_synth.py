"""The synthetic generator is test-data definition: pin its output."""
import hashlib

from text_alignment_b200 import synth


def test_appendix_c_input_digests(appendix_c):
    for rec in appendix_c:
        t, o = synth.make_pair(rec['seed'], rec['n'], rec['m'], rec['run_lo'], rec['run_hi'])
        assert len(t) == rec['n'] and len(o) == rec['m']
        assert hashlib.sha256((t + '\n' + o).encode()).hexdigest() == rec['input_sha256']


def test_config_shapes():
    t, o = synth.c2_pair(0)
    # gen_transcript may strip one trailing space, so len(t) is n or n-1
    assert 999 <= len(t) <= 1600 and abs(len(o) - 1.25 * len(t)) <= 2
    t, o = synth.c3_pair(5)
    assert 39 <= len(t) <= 120 and 40 <= len(o) <= 120
    t, o = synth.c4_pair(1)
    assert 599 <= len(t) <= 1000 and 2 * len(t) <= len(o) <= 4 * (len(t) + 1)
    for s in (t, o):
        assert '_' not in s and '~' not in s and '|' not in s


def test_bulk_numpy_layout():
    buf, t_off, n, o_off, m = synth.bulk_pairs_numpy(7, 16, 40, 120, lambda n, rng: rng.integers(40, 121))
    assert buf.dtype.name == 'uint8'
    assert int(n.sum() + m.sum()) == buf.size
    for k in range(16):
        assert o_off[k] == t_off[k] + n[k]
    assert 95 not in set(buf.tolist())      # no '_' in the alphabet
