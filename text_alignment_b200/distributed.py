"""One-process-per-GPU sharding of a packed batch.

Pairs are independent (the reference aligns one page per call, alignToOCR.py:273), so the
multi-GPU model is a partition: rank r aligns the r-th contiguous, cell-balanced range of
pairs on its own device; there is no data-path collective.  The only communication is a
host-side gather of the op strings / lengths / scores to rank 0 (torch.distributed object
gather over whatever backend the job initialised -- gloo on CPU hosts, NCCL under torchrun on
the GPU box), which is not part of the alignment itself.
"""
import numpy as np

from . import textSeqCompare as tsc


def shard_range(n, m, rank, world):
    """[lo, hi) of the pairs rank `rank` owns."""
    b = tsc.split_by_cells(np.asarray(n), np.asarray(m), world)
    return int(b[rank]), int(b[rank + 1])


def shard_batch(symbols, t_off, n, o_off, m, rank, world):
    """The rank's shard as a self-contained packed batch (symbols slice + rebased offsets)."""
    lo, hi = shard_range(n, m, rank, world)
    sub_sym, sub_t, sub_o = tsc._rebase(np.asarray(symbols), np.asarray(t_off)[lo:hi], np.asarray(n)[lo:hi],
                                        np.asarray(o_off)[lo:hi], np.asarray(m)[lo:hi])
    return (sub_sym, sub_t, np.asarray(n)[lo:hi], sub_o, np.asarray(m)[lo:hi]), (lo, hi)


def align_sharded(symbols, t_off, n, o_off, m, params, align_fn=None, group=None, device=None):
    """Collective call: every rank passes the SAME full batch description, aligns its own
    shard and rank 0 returns the gathered (ops, ops_off, ops_len, scores); other ranks return
    None.  `align_fn(symbols, t_off, n, o_off, m, params)` defaults to the device path on
    `device` (default: this rank's LOCAL_RANK)."""
    import os
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    shard, (lo, hi) = shard_batch(symbols, t_off, n, o_off, m, rank, world)
    if align_fn is None:
        dev = int(os.environ.get('LOCAL_RANK', '0')) if device is None else device

        def align_fn(s, t, nn, o, mm, p):
            return tsc.align_packed(s, t, nn, o, mm, p, devices=[dev])
    local = align_fn(*shard, params)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((lo, hi, local), gathered, dst=0, group=group)
    if rank != 0:
        return None
    gathered.sort(key=lambda g: g[0])
    bounds = np.array([g[0] for g in gathered] + [gathered[-1][1]], dtype=np.int64)
    return tsc.gather_shards([g[2] for g in gathered], np.asarray(n, dtype=np.int32),
                             np.asarray(m, dtype=np.int32), True, bounds)
