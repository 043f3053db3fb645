#!/bin/bash
# round 2, run V: generation 1 with the register-only copy path, L2 evict-first K/V streams, two-pass four-slot iterations, warp-per-tile logits
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py tests/test_pool_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py tests/test_exact_tokens_gpu.py tests/test_sampler_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert" | tail -12 > gpurun_out/v_tests.txt; cat gpurun_out/v_tests.txt
run() {  # name, chunks, env...
  local name=$1; shift; local chunks=$1; shift
  env "$@" WB200_MK_TRACE=gpurun_out/v_trace_$name.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 300 python bench.py --chunks $chunks --steps 2 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/v_bench_$name.json 2> gpurun_out/v_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/v_bench_$name.json').read().strip().splitlines()[-1])
    print('$name:', round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('$name: bench failed', e)
PY
  tail -1 gpurun_out/v_bench_$name.err
}
run g1_new 64 WB200_MK_PREFETCH=29
tail -24 gpurun_out/v_trace_g1_new.txt
run g1_noef 64 WB200_MK_PREFETCH=17
tail -3 gpurun_out/v_trace_g1_noef.txt
run g1_splitk_logits 64 WB200_MK_PREFETCH=13
tail -1 gpurun_out/v_trace_g1_splitk_logits.txt
