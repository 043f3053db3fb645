#!/bin/bash
# round 2, run N: encoder attention with 64-key blocks and two CTAs per SM
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py -q -m gpu -x -s 2>&1 | grep "passed\|failed\|FAILED\|worst\|Error\|error" | tail -8 > gpurun_out/n_kern.txt; cat gpurun_out/n_kern.txt
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert" | tail -30 > gpurun_out/n_e2e.txt; cat gpurun_out/n_e2e.txt
WB200_PROF_DUMP=gpurun_out/n_prof_dump.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
python - <<PY
import json
j=json.loads(open('gpurun_out/n_bench.json').read().strip().splitlines()[-1])
print(round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3), 'enc/window', round(j['engine']['encode_gpu_ms_per_window'],3), round(j['encode_roofline']['batched']['frac'],3), 'single', j['encode_ms'])
print(json.dumps(j['kernel_classes']))
PY
tail -1 gpurun_out/n_bench.err; sed -n 20,32p gpurun_out/n_prof_dump.txt
