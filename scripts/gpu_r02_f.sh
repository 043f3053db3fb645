#!/bin/bash
# round 2, call F: decode kernel with whole-pair cross-attention (head-major K/V): correctness + trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_e2e_gpu.py tests/test_golden_large_gpu.py tests/test_pool_gpu.py tests/test_kernels_gpu.py tests/test_exact_tokens_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | tail -150 > gpurun_out/f_pytest.txt
grep -n "prefix\|passed\|failed\|FAILED\|agreement\|rms" gpurun_out/f_pytest.txt | tail -40
WB200_MK_TRACE=gpurun_out/f_mk_trace.txt WB200_BENCH_REF_TOOL=0 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -22 gpurun_out/f_mk_trace.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['engine']['decode_gpu_ms_per_pass'], j['roofline']['frac'], j['encode_ms'], j['engine']['encode_gpu_ms_per_window'])
PY
