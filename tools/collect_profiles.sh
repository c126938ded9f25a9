#!/bin/bash
# Everything profiles/ is made of, in one gpurun call (one B200).  Every command first runs plain
# (exit code checked), then under ncu; numbers printed under ncu are never bench values.
#   tools/collect_profiles.sh <tag>      e.g. r2f  -> gpurun_out/<tag>_*
# Afterwards, here:  tools/ncu_summary.py gpurun_out/<tag>_pairs.ncu-rep profiles/<tag>_align_pairs_ncu.txt <cells> c2 10000  etc.
tag=${1:-r2}
out=gpurun_out
B="python bench.py --no-cpu-baseline --no-others --steps 2 --warmup 3"
set -x
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> /dev/null
for w in c3 c4 c5; do python bench.py --workload $w --no-cpu-baseline --no-others > $out/${tag}_bench_$w.json 2> /dev/null; done
# the launch list of the default command (its timed legs: warm-ups skipped by -s)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --no-cpu-baseline > /dev/null 2>&1
$B --workload c2 > $out/plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:align_pairs -s 4 -c 1 -f -o $out/${tag}_pairs $B --workload c2 > $out/ncu_c2.log 2>&1
$B --workload c3 > $out/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:align_lines16 -s 4 -c 1 -f -o $out/${tag}_lines16 $B --workload c3 > $out/ncu_c3.log 2>&1
$B --workload c5 > $out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:align_long -s 4 -c 1 -f -o $out/${tag}_long $B --workload c5 > $out/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_long -s 4 -c 1 -f -o $out/${tag}_trace_long $B --workload c5 > $out/ncu_c5t.log 2>&1
ls -la $out/${tag}_*
