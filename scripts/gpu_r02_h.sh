#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_e2e_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py tests/test_exact_tokens_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | tail -150 > gpurun_out/h_pytest.txt
grep -n "prefix\|passed\|failed\|FAILED\|identical history\|rms" gpurun_out/h_pytest.txt | tail -40
for v in 0 1; do
  if [ $v = 1 ]; then export WB200_GEMM_V1=1; fi
  WB200_BENCH_REF_TOOL=0 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/h_bench_v1_$v.json 2> gpurun_out/h_bench_v1_$v.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/h_bench_v1_$v.json').read().strip().splitlines()[-1])
print('GEMM_V1=$v', j['value'], j['e2e']['value'], j['engine']['decode_gpu_ms_per_pass'], j['roofline']['frac'], j['encode_ms'], j['encode_ms_parts'], j['engine']['encode_gpu_ms_per_window'], j['encode_roofline']['batched']['frac'], j['encode_roofline']['single_window']['frac'])
PY
  tail -2 gpurun_out/h_bench_v1_$v.err
done
