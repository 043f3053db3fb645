#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/j_gemm.txt; cat gpurun_out/j_gemm.txt
timeout -s KILL 900 python -m pytest tests/test_e2e_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py tests/test_kernels_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|TFLOP\|GB/s" | tail -20 > gpurun_out/j_e2e.txt; cat gpurun_out/j_e2e.txt
for c in 1 0; do
  WB200_GEMM_CLUSTER=$c WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/j_bench_cl$c.json 2> gpurun_out/j_bench_cl$c.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/j_bench_cl$c.json').read().strip().splitlines()[-1])
print('CLUSTER=$c', j['value'], j['e2e']['value'], j['engine']['decode_gpu_ms_per_pass'], j['roofline']['frac'], j['encode_ms'], j['encode_ms_parts'], j['engine']['encode_gpu_ms_per_window'], j['encode_roofline']['batched']['frac'], j['encode_roofline']['single_window']['frac'])
PY
  tail -2 gpurun_out/j_bench_cl$c.err
done
