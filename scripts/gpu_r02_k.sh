#!/bin/bash
# round 2, run K: row groups of the decode kernel on their own barriers (on / off, staggered starts), fast GELU + residual prefetch in the GEMM epilogue
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/k_gemm.txt; cat gpurun_out/k_gemm.txt
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py tests/test_pool_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py tests/test_exact_tokens_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert" | tail -30 > gpurun_out/k_e2e.txt; cat gpurun_out/k_e2e.txt
run() {  # tag, env...
  tag=$1; shift
  env "$@" WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/k_bench_$tag.json 2> gpurun_out/k_bench_$tag.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/k_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3), 'enc/window', round(j['engine']['encode_gpu_ms_per_window'],3), round(j['encode_roofline']['batched']['frac'],3), 'single', j['encode_ms'])
PY
  tail -1 gpurun_out/k_bench_$tag.err
}
run g0 WB200_MK_GROUPS=0
run g1 WB200_MK_GROUPS=1
run g1s20 WB200_MK_GROUPS=1 WB200_MK_STAGGER_US=20
run g1s50 WB200_MK_GROUPS=1 WB200_MK_STAGGER_US=50
run g1s100 WB200_MK_GROUPS=1 WB200_MK_STAGGER_US=100
WB200_MK_TRACE=gpurun_out/k_mk_trace.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ragged > /dev/null 2>&1; tail -30 gpurun_out/k_mk_trace.txt
