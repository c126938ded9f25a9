#!/usr/bin/env python
"""The steady loop of a profiled kernel as SASS, from an Nsight Compute report.

    python tools/sass_hot_loop.py gpurun_out/prof.ncu-rep profiles/r2_sass_pairs.txt

Takes the SASS view of the report's first kernel (`ncu --page source`, needs --import-source on),
keeps the instructions of the execution-count group that accounts for most executed instructions
-- the steady loop of the fill -- and writes them in program order with their execution counts and an opcode
histogram (SURVEY.md 8(d) "Evidence": VIADDMNMX / VIMNMX3 / SHFL.UP and no tensor or TMA ops)."""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else '?'
    hdr = rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    # the loop that accounts for most executed instructions: group by execution count (spin and
    # traceback loops run more often than the fill's steady loop, but are a few instructions long)
    groups = Counter()
    for r in data:
        groups[int(r[ix['Instructions Executed']] or 0)] += 1
    top = max(groups, key=lambda c: c * groups[c])
    hot = [r for r in data if 0.95 * top <= int(r[ix['Instructions Executed']] or 0) <= 1.05 * top]
    ops = Counter()
    for r in hot:
        tok = r[ix['Source']].split()
        op = tok[1] if tok[0].startswith('@') else tok[0]
        parts = op.split('.')
        ops['.'.join(parts[:2]) if parts[0] in ('VIADDMNMX', 'VIMNMX3', 'VIMNMX', 'SHFL', 'IMAD', 'LDG', 'STG') else parts[0]] += 1
    lines = ['# steady loop of %s' % kernel,
             '# %d instructions executed about %d times each (warp level), of %d in the kernel' % (len(hot), top, len(data)),
             '# opcode histogram of the loop:']
    lines += ['#   %-18s %d' % kv for kv in ops.most_common()]
    lines.append('')
    for r in hot:
        lines.append('%12s  %s' % (r[ix['Instructions Executed']], r[ix['Source']].strip()))
    open(dst, 'w').write('\n'.join(lines) + '\n')
    print('wrote', dst, len(hot), 'instructions')


if __name__ == '__main__':
    main()
