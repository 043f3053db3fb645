#!/bin/bash
# round 2, run AA: epilogue operands of the GEMV phases requested before the k-loop (prefetch bit 7): correctness + A/B
mkdir -p gpurun_out
WB200_MK_PREFETCH=189 timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py tests/test_pool_gpu.py -q -m gpu 2>&1 | tail -2
run() {
  local name=$1; shift
  env "$@" WB200_MK_TRACE=gpurun_out/aa_trace_$name.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/aa_bench_$name.json 2> gpurun_out/aa_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/aa_bench_$name.json').read().strip().splitlines()[-1])
    print('$name:', round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('$name: bench failed', e)
PY
  tail -1 gpurun_out/aa_bench_$name.err
}
run late WB200_MK_PREFETCH=61
run early WB200_MK_PREFETCH=189
tail -24 gpurun_out/aa_trace_early.txt | head -22
