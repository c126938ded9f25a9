"""In-process multi-device path: align_packed(devices=[0, 1, ...]) shards across the GPUs of one
box with one host thread per device and gathers on the host.  On a 1-GPU box the same code path
runs over two contexts on device 0 (devices=[0, 0])."""
import numpy as np
import pytest

from text_alignment_b200 import synth

pytestmark = pytest.mark.gpu


def test_two_devices_match_one_device():
    from text_alignment_b200 import _native, textSeqCompare as tsc
    pairs = [synth.c2_pair(k) for k in range(40)] + [synth.c3_pair(k) for k in range(500)] + [('', ''), ('a', '')]
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode(), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
    m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
    params = (8, -4, -7, -7, -3, 0, -1)
    one = tsc.align_packed(buf, t_off, n, t_off + n, m, params, devices=[0])
    devs = list(range(min(_native.device_count(), 8)))
    if len(devs) < 2:
        devs = [0, 0]            # two contexts (streams, arenas) on the one device
    many = tsc.align_packed(buf, t_off, n, t_off + n, m, params, devices=devs)
    assert np.array_equal(one[1], many[1]) and np.array_equal(one[2], many[2]) and np.array_equal(one[3], many[3])
    for k in range(len(pairs)):
        assert np.array_equal(one[0][one[1][k]:one[1][k] + one[2][k]], many[0][many[1][k]:many[1][k] + many[2][k]])


def _batch(pairs):
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode(), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
    m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
    return buf, t_off, n, t_off + n, m


def test_native_sharded_entry_bounds_and_errors():
    """tanw_align_batch_sharded: arbitrary shard boundaries (empty shards, one pair per shard),
    several contexts on one device, results in the caller's arrays; a shard's error comes back with
    its shard named; a context listed twice or a packed-ops context is refused."""
    from text_alignment_b200 import _native, textSeqCompare as tsc
    pairs = [synth.c3_pair(k) for k in range(300)] + [synth.c2_pair(k) for k in range(5)] + [('', 'abc'), ('', '')]
    buf, t_off, n, o_off, m = _batch(pairs)
    params = (8, -4, -7, -7, -3, 0, -1)
    P = len(pairs)
    one = tsc.align_packed(buf, t_off, n, o_off, m, params, devices=[0])
    ctxs = [tsc.get_context(0, replica=r) for r in range(3)]
    for bounds in ([0, 0, 300, P], [0, 1, 2, P], [0, P, P, P], [0, 150, 303, P]):
        ops_off, total = _native.Context.canonical_ops_layout(n, m)
        out = (np.full(total + 7, 9, np.uint8), np.zeros(P, np.int32), np.zeros((P, 3), np.int32))
        got = _native.Context.align_batch_sharded(ctxs, bounds, buf, t_off, n, o_off, m, params, out=out)
        assert got[0] is out[0]
        assert np.array_equal(one[2], got[2]) and np.array_equal(one[3], got[3])
        for k in range(P):
            assert np.array_equal(one[0][one[1][k]:one[1][k] + one[2][k]], got[0][got[1][k]:got[1][k] + got[2][k]])
        assert (out[0][total:] == 9).all()
    bad_n = n.copy(); bad_n[200] = -3
    with pytest.raises(ValueError) as e:               # TANW_E_INVALID, as from tanw_align_batch
        _native.Context.align_batch_sharded(ctxs, [0, 100, 250, P], buf, t_off, bad_n, o_off, m, params)
    assert 'shard 1' in str(e.value)
    with pytest.raises(ValueError):
        _native.Context.align_batch_sharded([ctxs[0], ctxs[0]], [0, 10, P], buf, t_off, n, o_off, m, params)
    with pytest.raises(ValueError):
        _native.Context.align_batch_sharded(ctxs, [0, 10, 5, P], buf, t_off, n, o_off, m, params)
    packed = _native.Context(0)
    packed.set_packed_ops(True)
    with pytest.raises(_native.NativeError):
        _native.Context.align_batch_sharded([ctxs[0], packed], [0, 10, P], buf, t_off, n, o_off, m, params)
    # the contexts are still usable
    again = _native.Context.align_batch_sharded(ctxs, [0, 100, 250, P], buf, t_off, n, o_off, m, params)
    assert np.array_equal(one[2], again[2])


def _sharded_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from text_alignment_b200 import _native, distributed as d, textSeqCompare as tsc
        pairs = [synth.c2_pair(k) for k in range(6)] + [synth.c3_pair(k) for k in range(60)]
        buf = np.frombuffer(''.join(t + o for t, o in pairs).encode(), dtype=np.uint8)
        n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
        m = np.array([len(o) for _, o in pairs], dtype=np.int32)
        t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
        params = (8, -4, -7, -7, -3, 0, -1)
        dev = rank % _native.device_count()              # one GPU per rank when there are enough
        out = d.align_sharded(buf, t_off, n, t_off + n, m, params, device=dev)
        if rank == 0:
            ref = tsc.align_packed(buf, t_off, n, t_off + n, m, params, devices=[0])
            ok = all(np.array_equal(out[i], ref[i]) for i in (1, 2, 3))
            ok = ok and all(np.array_equal(out[0][ref[1][k]:ref[1][k] + ref[2][k]], ref[0][ref[1][k]:ref[1][k] + ref[2][k]])
                            for k in range(len(pairs)))
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_one_process_per_gpu_sharding_with_host_gather():
    """distributed.align_sharded: every rank aligns its shard on its own device, rank 0 gathers
    on the host (gloo); no data-path collective.  Two ranks share GPU 0 on a 1-GPU box."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
