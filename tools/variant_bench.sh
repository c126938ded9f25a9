#!/bin/bash
# Run the parity tests and the headline bench once per kernel variant under build/variants/.
mkdir -p gpurun_out
for lib in build/variants/*.so; do
  name=$(basename $lib .so)
  echo "=== $name" >> gpurun_out/variants.log
  TANW_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not c5 and not 10k" 2>&1 | tail -1 >> gpurun_out/variants.log
  TANW_LIB=$PWD/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -1 > gpurun_out/bench_$name.json
  python - <<PY >> gpurun_out/variants.log
import json
try:
    d=json.loads(open('gpurun_out/bench_$name.json').read())
    print('$name', 'GCUPS', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'alu_frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$name', 'bench failed', e, open('gpurun_out/bench_$name.json').read()[-500:])
PY
done
cat gpurun_out/variants.log
