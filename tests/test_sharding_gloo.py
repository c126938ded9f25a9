"""World-size-2 test of the multi-GPU host logic on CPU (gloo): partition, per-rank shard
batches, host-side gather.  The per-shard aligner is the CPU oracle here (a stand-in for the
device so that the test needs no GPU); the real device path is exercised by
tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from text_alignment_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch():
    pairs = [synth.make_pair(100 + k, 5 + 7 * k % 90, 3 + 11 * k % 120, 1, 6) for k in range(23)]
    pairs[4] = ('', 'abc')
    pairs[9] = ('', '')
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode(), dtype=np.uint8)
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
    m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    t_off = np.concatenate([[0], np.cumsum(n.astype(np.int64) + m)[:-1]]).astype(np.int64)
    return buf, t_off, n, t_off + n, m


def _oracle_align(symbols, t_off, n, o_off, m, params):
    from oracle import nw_oracle
    sc, _ = nw_oracle.make_scoring(list(params[:6]), boundary_gap=params[6])
    ops, ops_off, ops_len, end3 = nw_oracle.align_batch_codes(symbols, t_off, n, o_off, m, sc, threads=1)
    scores = np.where(end3 <= -1e99, -1073741824, end3).astype(np.int32)
    return ops, ops_off, ops_len, scores


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from text_alignment_b200 import distributed as d
        batch = _batch()
        params = (8, -4, -7, -7, -3, 0, -1)
        out = d.align_sharded(*batch, params, align_fn=_oracle_align)
        if rank == 0:
            ref = _oracle_align(*batch, params)
            ok = all(np.array_equal(a, b) for a, b in zip(out, ref))
            lo, hi = d.shard_range(batch[2], batch[4], 0, world)
            q.put((ok, lo, hi, int(out[2].size)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_alignment_gathers_to_rank0(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ok, lo, hi, count = q.get(timeout=5)
    assert ok and lo == 0 and 0 < hi < 23 and count == 23


def test_shard_ranges_cover_everything():
    from text_alignment_b200 import distributed as d
    _, _, n, _, m = _batch()
    for world in (1, 2, 4, 8):
        edges = [d.shard_range(n, m, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n.size
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
