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
