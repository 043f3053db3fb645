#!/bin/bash
# round 2, run T: generation 2 with LayerNorm parameters through shared memory and a register-only copy path in the cross-attention; generation 1 with the same copy path
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | grep "passed\|failed\|FAILED\|Error\|error\|assert" | tail -5 > gpurun_out/t_tests.txt; cat gpurun_out/t_tests.txt
run() {  # name, chunks, env...
  local name=$1; shift; local chunks=$1; shift
  env "$@" WB200_MK_TRACE=gpurun_out/t_trace_$name.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 300 python bench.py --chunks $chunks --steps 1 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/t_bench_$name.json 2> gpurun_out/t_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/t_bench_$name.json').read().strip().splitlines()[-1])
    print('$name:', round(j['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('$name: bench failed', e)
PY
  tail -1 gpurun_out/t_bench_$name.err
}
run g2_64 64 WB200_MK_GEN=2
tail -34 gpurun_out/t_trace_g2_64.txt
run g1_64 64 WB200_MK_GEN=1
tail -24 gpurun_out/t_trace_g1_64.txt | head -14
run g2_64_st0 64 WB200_MK_GEN=2 WB200_MK_STAGGER=0
run g2_16 16 WB200_MK_GEN=2
tail -34 gpurun_out/t_trace_g2_16.txt | head -22
