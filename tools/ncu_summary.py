#!/usr/bin/env python
"""Summarise an Nsight Compute report into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_align_pairs.txt [cells] [workload pairs]

With `workload pairs` (e.g. `c2 10000`) the DRAM bytes of the first profiled launch are also
recorded in profiles/traffic.json together with a digest of the kernel sources, which is where
bench.py takes `roofline.traffic` from (and marks it stale once the kernels change).

Keeps the metrics the roofline discussion in DESIGN.md refers to, the stall breakdown and the
executed-instruction mix per opcode (from the source page; needs -lineinfo / --import-source)."""
import csv
import io
import subprocess
import sys
from collections import Counter

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']


def ncu(page, rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    cells = float(sys.argv[3]) if len(sys.argv) > 3 else None
    lines = ['# summary of %s (ncu --set full --clock-control none --import-source on)' % rep.split('/')[-1]]
    rows = ncu('raw', rep)
    hdr, units = rows[0], rows[1]
    for li, vals in enumerate(rows[2:]):
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        lines.append('')
        lines.append('## launch %d: %s' % (li, name))
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                lines.append('%-80s %-14s %s' % (h, u, v))
    src = ncu('source', rep)
    if len(src) > 2:
        h = src[1]
        ix = {k: i for i, k in enumerate(h)}
        data = [r for r in src[2:] if len(r) == len(h)]
        c = Counter()
        for r in data:
            tok = r[ix['Source']].split()
            op = tok[1] if tok[0].startswith('@') else tok[0]
            c[op.split('.')[0]] += int(r[ix['Instructions Executed']] or 0)
        tot = sum(c.values())
        lines.append('')
        lines.append('## executed warp-instructions by opcode (first profiled launch, source page)')
        lines.append('total %d%s' % (tot, '  = %.2f thread-instr per cell' % (tot * 32 / cells) if cells else ''))
        for k, v in c.most_common(24):
            lines.append('%-12s %14d %s' % (k, v, '%.3f /cell' % (v * 32 / cells) if cells else ''))
        lines.append('')
        lines.append('## top stall sites (samples)')
        for col in ('stall_long_sb', 'stall_wait', 'stall_math', 'stall_branch_resolving', 'stall_short_sb'):
            if col not in ix:
                continue
            tot_s = sum(int(r[ix[col]] or 0) for r in data)
            lines.append('%s total %d' % (col, tot_s))
            for r in sorted(data, key=lambda r: -int(r[ix[col]] or 0))[:4]:
                lines.append('    %8s  %s' % (r[ix[col]], r[ix['Source']].strip()[:90]))
    open(dst, 'w').write('\n'.join(lines) + '\n')
    print('wrote', dst)
    if len(sys.argv) > 5:
        import json
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import bench
        first = rows[2]
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        total = 0.0
        for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(name)
            total += float(first[i].replace(',', '')) * scale.get(units[i], 1.0)
        path = os.path.join(root, 'profiles', 'traffic.json')
        try:
            tab = json.load(open(path))
        except (OSError, ValueError):
            tab = {}
        tab[sys.argv[4]] = dict(pairs=int(sys.argv[5]), dram_bytes_per_launch=total,
                                kernel=first[hdr.index('Kernel Name')], digest=bench.source_digest(),
                                source=os.path.relpath(dst, root))
        json.dump(tab, open(path, 'w'), indent=1, sort_keys=True)
        print('updated', path)


if __name__ == '__main__':
    main()
