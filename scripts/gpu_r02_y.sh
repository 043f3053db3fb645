#!/bin/bash
# round 2, run Y: barrier variant (every CTA polls the arrival counter) A/B; plugin with two host states
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_zz_ggml_backend_plugin_gpu.py -q -m gpu -s -k two_host 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | tail -6 > gpurun_out/y_plugin.txt; cat gpurun_out/y_plugin.txt
WB200_MK_PREFETCH=61 timeout -s KILL 600 python -m pytest tests/test_e2e_gpu.py tests/test_pool_gpu.py -q -m gpu 2>&1 | tail -2
run() {
  local name=$1; shift
  env "$@" WB200_MK_TRACE=gpurun_out/y_trace_$name.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/y_bench_$name.json 2> gpurun_out/y_bench_$name.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/y_bench_$name.json').read().strip().splitlines()[-1])
    print('$name:', round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('$name: bench failed', e)
PY
  tail -1 gpurun_out/y_bench_$name.err
}
run flags WB200_MK_PREFETCH=29
run poll WB200_MK_PREFETCH=61
tail -24 gpurun_out/y_trace_poll.txt | head -14
