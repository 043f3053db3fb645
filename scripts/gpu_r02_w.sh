#!/bin/bash
# round 2, run W: beam search with the categorical draws on the device (k candidates per decoder come back instead of n_vocab floats)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_sampler_gpu.py tests/test_exact_tokens_gpu.py tests/test_e2e_gpu.py tests/test_pool_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|Error\|error\|assert\|beam\|identical" | tail -20 > gpurun_out/w_tests.txt; cat gpurun_out/w_tests.txt
# config 3 (large-v3 Q4_K, beam 5, 30-minute clip) with device draws and with the host sampler
for hb in 0 1; do
  if [ $hb = 1 ]; then export WB200_HOST_BEAM=1; else unset WB200_HOST_BEAM; fi
  WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --config 3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/w_cfg3_hostbeam$hb.json 2> gpurun_out/w_cfg3_hostbeam$hb.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/w_cfg3_hostbeam$hb.json').read().strip().splitlines()[-1])
    print('config 3 host_beam=$hb:', {k: j[k] for k in ('value','unit','ms_per_step') if k in j}, j.get('engine', {}))
except Exception as e: print('config 3 failed', e)
PY
  tail -2 gpurun_out/w_cfg3_hostbeam$hb.err
done
