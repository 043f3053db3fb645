#!/bin/bash
# round 2, run P: state after the container was re-created: core GPU tests, phase trace, default-ish bench line
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/test_e2e_gpu.py tests/test_pool_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py tests/test_exact_tokens_gpu.py tests/test_kernels_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert" | tail -30 > gpurun_out/p_e2e.txt; cat gpurun_out/p_e2e.txt
WB200_MK_TRACE=gpurun_out/p_mk_trace.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
python - <<PY
import json
j=json.loads(open('gpurun_out/p_bench.json').read().strip().splitlines()[-1])
print(round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3), 'enc/window', round(j['engine']['encode_gpu_ms_per_window'],3), round(j['encode_roofline']['batched']['frac'],3), 'single', j['encode_ms'])
PY
tail -1 gpurun_out/p_bench.err; tail -40 gpurun_out/p_mk_trace.txt
