#!/bin/bash
# round 2, call E: A/B of decode-kernel variants (phase traces), then correctness of the newest one
mkdir -p gpurun_out
for v in v0 v2 v3; do
  WB200_LIB=$PWD/whisper.cpp_b200/libwhisper_b200_$v.so WB200_MK_TRACE=gpurun_out/e_trace_$v.txt WB200_BENCH_REF_TOOL=0 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/e_bench_$v.json 2> gpurun_out/e_bench_$v.err
  echo "== $v"; tail -22 gpurun_out/e_trace_$v.txt | head -14; tail -8 gpurun_out/e_trace_$v.txt
  python - <<PY
import json
j=json.loads(open('gpurun_out/e_bench_$v.json').read().strip().splitlines()[-1])
print('$v', j['value'], j['engine']['decode_gpu_ms_per_pass'], j['encode_ms'], j['engine']['encode_gpu_ms_per_window'])
PY
done
timeout 1200 python -m pytest tests/test_e2e_gpu.py tests/test_golden_large_gpu.py tests/test_pool_gpu.py tests/test_kernels_gpu.py tests/test_exact_tokens_gpu.py -q -m gpu 2>&1 | grep -v "^whisper_\|^ggml_" | tail -40 > gpurun_out/e_pytest.txt
tail -12 gpurun_out/e_pytest.txt
