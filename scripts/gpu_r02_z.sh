#!/bin/bash
# round 2, run Z: LayerNorm folded into the consumer's row staging in generation 1 (prefetch bit 6): correctness + A/B
mkdir -p gpurun_out
WB200_MK_PREFETCH=125 timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py tests/test_pool_gpu.py tests/test_kernels_gpu.py tests/test_golden_large_gpu.py tests/test_exact_tokens_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert\|identical prefix" | tail -16 > gpurun_out/z_tests.txt; cat gpurun_out/z_tests.txt
run() {
  local name=$1; shift
  env "$@" WB200_MK_TRACE=gpurun_out/z_trace_$name.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/z_bench_$name.json 2> gpurun_out/z_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/z_bench_$name.json').read().strip().splitlines()[-1])
    print('$name:', round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('$name: bench failed', e)
PY
  tail -1 gpurun_out/z_bench_$name.err
}
run fold WB200_MK_PREFETCH=125
tail -24 gpurun_out/z_trace_fold.txt
run nofold WB200_MK_PREFETCH=61
