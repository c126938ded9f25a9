#!/usr/bin/env python
"""bench.py -- headline benchmark of the affine-gap NW hot path (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4] [--pairs P]

A "step" is one pass of the hot path over one batch: `--pairs` synthetic pairs of the chosen
BASELINE config per GPU (default config 2: 10 000 seeded page pairs of 1-2k characters;
SURVEY.md 8(d)), default scoring [8,-4,-7,-7,-3,0].  Weak scaling: every rank aligns its own
batch, no data-path collective (pairs are independent).

  value  : whole-job GCUPS (sum over ranks of n*m / max-over-ranks device time), inputs
           resident in HBM, timed with CUDA events on the library's stream;
  e2e    : the same metric through the C-ABI call tanw_align_batch with pinned HOST buffers:
           H2D of symbols + pair table, fill, traceback, D2H of op strings / lengths / scores
           inside the timed region;
  roofline / cpu_baseline : see DESIGN.md "Measurement".

`--impl reference` times the CPU restatement of the reference's pure-Python aligner
(oracle/py_port.py; the reference itself, being Python, cannot travel to the GPU box) over
all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'batched affine-NW GCUPS'      # BASELINE.json metric; pages/sec is reported beside it as pages_per_s
OPS_PER_CELL = 21          # algorithmic int32 ops per cell of the reference recurrence
PTR_BYTES_PER_CELL = 1     # three 2-bit pointers packed in one byte
DEFAULT_PARAMS = (8, -4, -7, -7, -3, 0, -1)
NOMINAL_INT32_PEAK = 148 * 64 * 1.965e9     # alu pipe, lane-ops/s


def ncu_traffic(workload, npairs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed ncu --set full summary of the same workload (profiles/); None if it does not apply."""
    if workload != 'c2' or npairs != WORKLOADS['c2']['default_pairs']:
        return None
    path = os.path.join(ROOT, 'profiles', 'r1m_align_pairs_ncu.txt')
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    total = 0.0
    try:
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                total += float(f[2]) * scale.get(f[1], 1.0)
    except (OSError, ValueError):
        return None
    return total or None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get('hbm_gbs', 6650.0)), 'measured'
    return 6650.0, 'fallback'


# ---- workload ---------------------------------------------------------------------------------

def _gen_one(args):
    from text_alignment_b200 import synth
    which, k = args
    if which == 'c1':
        return synth.c1_page()
    if which == 'c5':
        return synth.c5_pair()
    return getattr(synth, which + '_pair')(k)


def make_workload(which, first, count, procs):
    """Seeded pairs `first .. first+count` of a BASELINE config, packed for the C ABI."""
    jobs = [(which, first + k) for k in range(count)]
    if procs > 1 and count >= 64:
        with mp.get_context('fork').Pool(procs) as pool:
            pairs = pool.map(_gen_one, jobs, chunksize=max(1, count // (procs * 8)))
    else:
        pairs = [_gen_one(j) for j in jobs]
    return pack_pairs(pairs), pairs


def pack_pairs(pairs):
    n = np.array([len(t) for t, _ in pairs], dtype=np.int32)
    m = np.array([len(o) for _, o in pairs], dtype=np.int32)
    buf = np.frombuffer(''.join(t + o for t, o in pairs).encode('latin-1'), dtype=np.uint8).copy()
    lens = n.astype(np.int64) + m
    t_off = np.zeros(len(pairs), dtype=np.int64)
    if len(pairs):
        np.cumsum(lens[:-1], out=t_off[1:])
    return buf, t_off, n, t_off + n, m


WORKLOADS = {
    'c2': dict(name='config 2: seeded synthetic page pairs, n~U[1000,1600], m=1.25n, 20% sub + 5% indel, runs 5-40',
               default_pairs=10000),
    'c1': dict(name='config 1: single Salzinnes-shaped page, seed 1001, n=1200, m=1500', default_pairs=1),
    'c3': dict(name='config 3: seeded synthetic line pairs, n,m~U[40,120], runs 2-6 (1M pairs over 8 GPUs)',
               default_pairs=125000),
    'c4': dict(name='config 4: St. Gall-shaped pages, n~U[600,1000], m=n*U[2,4], inserted runs 50-400',
               default_pairs=4096),
    'c5': dict(name='config 5: whole-manuscript pair, seed 5001, n=80000 x m=100000 (chained-pass path)',
               default_pairs=1),
}


# ---- clocks ----------------------------------------------------------------------------------

class ClockSampler(object):
    """SM clock / throttle-reason sampling DURING the timed region (profiling recipe's clocks
    line), through NVML in a background thread (an nvidia-smi -lms child polling the driver
    was measured to stall synchronous CUDA calls of the timed process by several ms)."""

    def __init__(self, gpu_index, period_s=0.05):
        self.gpu = gpu_index
        self.period = period_s
        self.samples = []
        self.stop_flag = threading.Event()
        self.err = None
        self.th = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            idx = self.gpu
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                try:
                    idx = int(vis.split(',')[self.gpu])
                except (ValueError, IndexError):
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                       # noqa: BLE001
            self.err = repr(e)
            return
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:                    # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), clk, reasons, pw))
            except Exception as e:                   # noqa: BLE001
                self.err = repr(e)
            self.stop_flag.wait(self.period)

    def stop(self, t0, t1):
        self.stop_flag.set()
        if self.th is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvml unavailable: %s' % self.err])
        self.th.join(timeout=1.0)
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in inside))
        return dict(sm_mhz=float(np.median([s[1] for s in inside])) if inside else None,
                    sm_max_mhz=float(self.smax), power_w_max=max(s[3] for s in inside) if inside else None,
                    samples=len(inside), reasons=reasons, how='NVML, %d ms period' % int(self.period * 1e3))


# ---- CPU baseline (oracle: the only place bench.py may execute oracle/) -------------------------

def _py_port_one(args):
    from oracle import py_port
    t, o = args
    t0 = time.perf_counter()
    py_port.perform_alignment(list(t), list(o))
    return len(t) * len(o), time.perf_counter() - t0


def crop_pairs(pairs, cells_per_pair):
    out = []
    for t, o in pairs:
        f = min(1.0, (cells_per_pair / max(1.0, float(len(t)) * len(o))) ** 0.5)
        out.append((t[:max(1, int(len(t) * f))], o[:max(1, int(len(o) * f))]))
    return out


def cpu_python_port(pairs, cores, target_s):
    """The reference's algorithm in pure Python (oracle/py_port.py, ~10 us per cell like the
    reference), one pair per host core, pairs cropped so a step lasts about target_s."""
    sample = crop_pairs(pairs[:cores], target_s / 10e-6)
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(_py_port_one, sample, chunksize=1)
    wall = time.perf_counter() - t0
    cells = sum(c for c, _ in res)
    return cells, wall, sample


def cpu_c_oracle(packed, cores, max_pairs):
    from oracle import nw_oracle
    buf, t_off, n, o_off, m = packed
    k = min(max_pairs, n.size)
    sc, _ = nw_oracle.make_scoring(list(DEFAULT_PARAMS[:6]), boundary_gap=DEFAULT_PARAMS[6])
    t0 = time.perf_counter()
    nw_oracle.align_batch_codes(buf, t_off[:k], n[:k], o_off[:k], m[:k], sc, threads=cores, want_scores=False)
    wall = time.perf_counter() - t0
    cells = int((n[:k].astype(np.int64) * m[:k]).sum())
    return cells, wall, k


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    wl = WORKLOADS[args.workload]
    _, pairs = make_workload(args.workload, 0, cores, min(cores, 16))
    total = args.steps + args.warmup
    target_s = max(0.5, min(20.0, 150.0 / max(total, 1)))
    for _ in range(args.warmup):
        cpu_python_port(pairs, cores, target_s)
    cells = 0
    wall = 0.0
    sample = None
    for _ in range(args.steps):
        c, w, sample = cpu_python_port(pairs, cores, target_s)
        cells += c
        wall += w
    gcups = cells / wall / 1e9
    desc = ('%d crops of %s pages per step (first ~%dx%d chars of seeds 2000000..), one per host core, '
            'pure-Python restatement of textSeqCompare.py (oracle/py_port.py); the reference is Python and '
            'cannot travel to the GPU box' % (len(sample), args.workload, len(sample[0][0]), len(sample[0][1])))
    line = dict(impl='reference', metric=METRIC, value=gcups, unit='GCUPS', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=wall / max(args.steps, 1) * 1e3,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64 (python float)', data='synthetic',
                config=dict(workload=wl['name'], scoring=list(DEFAULT_PARAMS[:6])),
                pages_per_s=len(sample) * args.steps / wall,
                cpu_baseline=dict(value=gcups, unit='GCUPS', cores=cores, kind='port', sample=desc),
                e2e=dict(value=gcups, unit='GCUPS', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))
    return 0


# ---- our arm ---------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--out', default='', help='also write the JSON line to this file')
    ap.add_argument('--pairs', type=int, default=0, help='pairs per GPU per step (default: the config size)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parity-pairs', type=int, default=16)
    ap.add_argument('--band-rows', type=int, default=0,
                    help='cut chained-stripe pairs (c5, c1) into row bands of this height (checkpoint + recompute)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from text_alignment_b200 import _native
    from text_alignment_b200.textSeqCompare import get_context

    wl = WORKLOADS[args.workload]
    npairs = args.pairs or wl['default_pairs']
    cores = len(os.sched_getaffinity(0))
    gen_procs = max(1, min(32, cores // max(world, 1)))
    packed, pairs = make_workload(args.workload, rank * npairs, npairs, gen_procs)
    buf, t_off, n, o_off, m = packed
    cells = int((n.astype(np.int64) * m).sum())

    ctx = get_context(local_rank)
    ctx.set_long_band_rows(args.band_rows)
    scoring = ctx.make_scoring(*DEFAULT_PARAMS)

    # ---- parity spot check against the oracle (outside every timed region) ----------------------
    parity = 0
    if args.workload == 'c5':
        args.parity_pairs = 0      # 8e9 cells: checked by tests/test_gpu_parity.py::test_c5_whole_manuscript_pair
    if args.parity_pairs > 0:
        from oracle import nw_oracle
        k = min(args.parity_pairs, npairs)
        sub = pack_pairs(pairs[:k])
        got = ctx.align_batch(*sub, scoring)
        sc, _ = nw_oracle.make_scoring(list(DEFAULT_PARAMS[:6]), boundary_gap=DEFAULT_PARAMS[6])
        want = nw_oracle.align_batch_codes(*sub, sc, threads=min(cores, 16))
        ok = np.array_equal(got[2], want[2]) and all(
            np.array_equal(got[0][got[1][i]:got[1][i] + got[2][i]], want[0][want[1][i]:want[1][i] + want[2][i]])
            for i in range(k))
        ok = ok and np.array_equal(got[3].astype(np.float64), np.where(want[3] <= -1e99, _native.NEG_INF, want[3]))
        if not ok:
            print(json.dumps(dict(error='parity check against the oracle FAILED; no number reported')))
            return 2
        parity = k

    # ---- pinned host buffers for the end-to-end leg ------------------------------------------------
    def pinned(a):
        t = torch.empty(max(a.size, 1), dtype=torch.uint8, pin_memory=True)
        v = t.numpy()[:a.size]
        v[...] = a
        return t, v
    keep = []
    p_buf = pinned(buf); keep.append(p_buf)
    ops_cap = int((n.astype(np.int64) + m).sum())
    t_ops = torch.empty(max(ops_cap, 1), dtype=torch.uint8, pin_memory=True)
    t_len = torch.empty(max(npairs, 1), dtype=torch.int32, pin_memory=True)
    t_sc = torch.empty((max(npairs, 1), 3), dtype=torch.int32, pin_memory=True)
    out = (t_ops.numpy(), t_len.numpy(), t_sc.numpy())

    ext = torch.cuda.ExternalStream(ctx.stream_handle(), device=torch.device('cuda', local_rank))
    # measured issue rates (warp-lane instructions/s): IADD3, VIMNMX, fused VIADDMNMX
    # and VIADD + LOP3 together (both integer pipes busy: the issue ceiling of a kernel that mixes them)
    rate_add, rate_max, rate_fused, rate_two = (ctx.measure_int32_peak(w) for w in (0, 1, 2, 3))
    alu_pipe_peak = max(rate_add, rate_max)
    int32_peak = max(rate_two, alu_pipe_peak)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: inputs resident in HBM, K launches back to back, CUDA events on the lib stream ----
    ctx.prepare(p_buf[1], t_off, n, o_off, m, scoring)
    for _ in range(args.warmup):
        ctx.run()
    ctx.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record(ext)
    for _ in range(args.steps):
        ctx.run()
    e1.record(ext)
    ctx.sync()
    barrier()
    w1 = time.perf_counter()
    dev_ms = e0.elapsed_time(e1)
    launches = args.steps * ctx.timing()['kernel_launches']

    # ---- leg 2: end to end through the C ABI with host buffers ------------------------------------
    for _ in range(max(args.warmup, 5)):      # first calls on fresh pinned buffers are erratic (20-100 ms)
        ctx.align_batch(p_buf[1], t_off, n, o_off, m, scoring, out=out)
    barrier()
    x0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        s0 = time.perf_counter()
        ctx.align_batch(p_buf[1], t_off, n, o_off, m, scoring, out=out)
        step_ms.append(round((time.perf_counter() - s0) * 1e3, 3))
    barrier()
    x1 = time.perf_counter()
    tm = ctx.timing()
    e2e_s = x1 - x0

    # ---- leg 3: the same end-to-end steps, double buffered over two contexts ----------------------
    # A streaming caller overlaps step k+1's upload and step k-1's download with step k's kernel
    # by alternating two contexts (each owns a stream, device buffers and a pointer arena).  Every
    # step still uploads its inputs from pinned host memory and downloads its results.
    ctx_b = _native.Context(local_rank)
    ctx_b.set_long_band_rows(args.band_rows)
    t_ops2 = torch.empty(max(ops_cap, 1), dtype=torch.uint8, pin_memory=True)
    t_len2 = torch.empty(max(npairs, 1), dtype=torch.int32, pin_memory=True)
    t_sc2 = torch.empty((max(npairs, 1), 3), dtype=torch.int32, pin_memory=True)
    lanes = [(ctx, out), (ctx_b, (t_ops2.numpy(), t_len2.numpy(), t_sc2.numpy()))]

    def fetch_into(c, o):
        ops_off, total = c.canonical_ops_layout(n, m)
        c._check(c._lib.tanw_batch_fetch(c._h, _native._ptr(o[0], _native._u8p), _native._ptr(ops_off, _native._i64p),
                                         o[0].size, _native._ptr(o[1], _native._i32p), _native._ptr(o[2], _native._i32p)))

    def pipelined(steps):
        pending = []
        for k in range(steps):
            c, o = lanes[k % 2]
            if len(pending) == 2:
                fetch_into(*pending.pop(0))
            c.prepare(p_buf[1], t_off, n, o_off, m, c.make_scoring(*DEFAULT_PARAMS))
            c.run()
            pending.append((c, o))
        for c, o in pending:
            fetch_into(c, o)
    pipelined(4)
    barrier()
    y0 = time.perf_counter()
    pipelined(args.steps)
    barrier()
    y1 = time.perf_counter()
    pipe_s = y1 - y0
    same = bool(np.array_equal(out[0][:ops_cap], t_ops2.numpy()[:ops_cap]))
    ctx_b.close()
    clocks = sampler.stop(w0, y1)

    # ---- the Python list <-> buffer shim, reported separately (it is not the path) ----
    shim = None
    if rank == 0 and args.workload in ('c2', 'c4', 'c1'):
        from text_alignment_b200 import textSeqCompare as tsc_mod
        k = min(256, npairs)
        lists = [(list(t), list(o)) for t, o in pairs[:k]]
        tsc_mod.perform_alignment_batch(lists[:8], devices=[local_rank])
        s0 = time.perf_counter()
        tsc_mod.perform_alignment_batch(lists, devices=[local_rank])
        s1 = time.perf_counter()
        shim = dict(pages=k, ms_per_page=(s1 - s0) * 1e3 / k, pages_per_s=k / (s1 - s0),
                    what='perform_alignment_batch on Python lists: interning to uint8 codes, one launch, '
                         'op strings back to two lists per page')

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s * 1e3, w1 - w0, pipe_s * 1e3], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, wall_s, pipe_ms = t.tolist()
        c = torch.tensor([cells, npairs], dtype=torch.float64, device='cuda')
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_cells, tot_pairs = c.tolist()
    else:
        e2e_ms, wall_s, tot_cells, tot_pairs = e2e_s * 1e3, w1 - w0, float(cells), float(npairs)
        pipe_ms = pipe_s * 1e3

    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        gcups = tot_cells * args.steps / (dev_ms * 1e-3) / 1e9
        e2e_gcups = tot_cells * args.steps / (e2e_ms * 1e-3) / 1e9
        launch_s = dev_ms * 1e-3 / args.steps                     # one launch per step on this rank
        ach_ops = cells * OPS_PER_CELL / launch_s
        ach_gbs = cells * PTR_BYTES_PER_CELL / launch_s / 1e9
        line = dict(
            metric=METRIC, value=gcups, unit='GCUPS', n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='weak',
            vs_baseline=None, dtype='int32', data='synthetic',
            config=dict(workload=wl['name'], pairs_per_gpu=npairs, cells_per_gpu=cells,
                        scoring=list(DEFAULT_PARAMS[:6]), band_rows=args.band_rows, parallelism='pairs sharded, %d rank(s)' % world,
                        l2='every step writes %d MB of traceback pointers per GPU (>> 126 MB L2)' % (cells // 2 ** 20),
                        parity_checked_pairs=parity),
            pages_per_s=tot_pairs * args.steps / (dev_ms * 1e-3),
            wall_ms_per_step=wall_s * 1e3 / args.steps,
            e2e=dict(value=e2e_gcups, unit='GCUPS', h2d_bytes_per_step=int(tm['h2d_bytes']),
                     d2h_bytes_per_step=int(tm['d2h_bytes']), pages_per_s=tot_pairs * args.steps / (e2e_ms * 1e-3),
                     ms_per_step=e2e_ms / args.steps, rank0_step_ms=step_ms,
                     double_buffered=dict(value=tot_cells * args.steps / (pipe_ms * 1e-3) / 1e9, unit='GCUPS',
                                          ms_per_step=pipe_ms / args.steps, results_identical=same,
                                          how='two contexts alternate; each step = prepare (H2D) + run + fetch (D2H)'),
                     breakdown_ms=dict(h2d=tm['h2d_ms'], kernel=tm['kernel_ms'], d2h=tm['d2h_ms'])),
            gpu_launches=launches,
            roofline=dict(bound='alu', achieved=ach_ops / 1e12, peak=int32_peak / 1e12, unit='Tlane-op/s',
                          frac=ach_ops / int32_peak, traffic=ncu_traffic(args.workload, npairs),
                          note='achieved = %d algorithmic int32 ops/cell (SURVEY 8(d)) x cells / launch time; '
                               'peak = measured dependency-free issue rate of the alu and fma integer pipes '
                               'together (VIADD + LOP3 alternating, tanw_measure_int32_peak(3)): strip_row puts '
                               'its IMAD/VIADD work on the fma pipe, so the single alu pipe (SURVEY 8(d): %.1f '
                               'nominal) is not its ceiling -- see frac_of_alu_pipe'
                               % (OPS_PER_CELL, NOMINAL_INT32_PEAK / 1e12),
                          measured_rates=dict(iadd3=rate_add / 1e12, vimnmx=rate_max / 1e12,
                                              viaddmnmx=rate_fused / 1e12, viadd_plus_lop3=rate_two / 1e12),
                          frac_of_alu_pipe=ach_ops / alu_pipe_peak,
                          frac_of_nominal_alu_pipe=ach_ops / NOMINAL_INT32_PEAK,
                          hbm=dict(bound='hbm', achieved=ach_gbs, peak=hbm_peak, unit='GB/s',
                                   frac=ach_gbs / hbm_peak, peak_source=hbm_src)),
            clocks=clocks)
        if shim:
            line['python_list_shim'] = shim
        if not args.no_cpu_baseline:
            t_s = 12.0
            c_cells, c_wall, sample = cpu_python_port(pairs, cores, t_s)
            line['cpu_baseline'] = dict(
                value=c_cells / c_wall / 1e9, unit='GCUPS', cores=cores, kind='port',
                sample='%d crops (~%dx%d chars) of the same pages, one per host core, pure-Python restatement '
                       'of textSeqCompare.py (oracle/py_port.py)' % (len(sample), len(sample[0][0]), len(sample[0][1])))
            k_cells, k_wall, k = cpu_c_oracle(packed, cores, max(cores * 4, 64))
            line['cpu_baseline_c'] = dict(value=k_cells / k_wall / 1e9, unit='GCUPS', cores=cores, kind='port',
                                          sample='%d full pages, oracle/nw_oracle.c (scalar C, float64), %d threads' % (k, cores))
        print(json.dumps(line))
        if args.out:
            with open(args.out, 'w') as f:
                json.dump(line, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
