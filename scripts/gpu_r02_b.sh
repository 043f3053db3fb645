#!/bin/bash
# round 2, call B: large-shape goldens, exact-token tests, pool tests, bench through whisper.h
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_golden_large_gpu.py tests/test_exact_tokens_gpu.py tests/test_pool_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_" | tail -60 > gpurun_out/b_pytest.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -40 gpurun_out/b_pytest.txt; python - <<'PY'
import json
j=json.loads(open('gpurun_out/b_bench.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'], j['e2e']['d2h_bytes_per_step'], j['e2e_threads']['value'])
PY
tail -5 gpurun_out/b_bench.err
