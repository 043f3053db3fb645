#!/bin/bash
# round 2, call D: new mk_gemv -- correctness (persistent kernel vs chain vs reference, large-shape goldens, pool) + phase trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_e2e_gpu.py tests/test_golden_large_gpu.py tests/test_pool_gpu.py tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | grep -v "^whisper_\|^ggml_" | tail -30 > gpurun_out/d_pytest.txt
tail -5 gpurun_out/d_pytest.txt
WB200_MK_TRACE=gpurun_out/d_mk_trace.txt timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -24 gpurun_out/d_mk_trace.txt
python - <<'PY'
import json
j=json.loads(open('gpurun_out/d_bench.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'], j['engine']['decode_gpu_ms_per_pass'], j['roofline']['frac'])
PY
tail -3 gpurun_out/d_bench.err
