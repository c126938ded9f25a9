"""In-process multi-device path: align_packed(devices=[0, 1, ...]) shards across the GPUs of one
box with one host thread per device and gathers on the host.  Skipped on a 1-GPU box."""
import numpy as np
import pytest

from text_alignment_b200 import synth

pytestmark = pytest.mark.gpu


def test_two_devices_match_one_device():
    from text_alignment_b200 import _native, textSeqCompare as tsc
    if _native.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    pairs = [synth.c2_pair(k) for k in range(40)] + [synth.c3_pair(k) for k in range(500)] + [('', ''), ('a', '')]
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode(), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
    m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
    params = (8, -4, -7, -7, -3, 0, -1)
    one = tsc.align_packed(buf, t_off, n, t_off + n, m, params, devices=[0])
    devs = list(range(min(_native.device_count(), 8)))
    many = tsc.align_packed(buf, t_off, n, t_off + n, m, params, devices=devs)
    assert np.array_equal(one[1], many[1]) and np.array_equal(one[2], many[2]) and np.array_equal(one[3], many[3])
    for k in range(len(pairs)):
        assert np.array_equal(one[0][one[1][k]:one[1][k] + one[2][k]], many[0][many[1][k]:many[1][k] + many[2][k]])
