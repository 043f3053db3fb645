#!/bin/bash
# round 2, run Q: second-generation decode kernel (row groups, 2 CTAs per SM): correctness first, then trace + bench, A/B against generation 1
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py tests/test_pool_gpu.py -q -m gpu -x -s 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert\|decode" | tail -30 > gpurun_out/q_tests.txt; cat gpurun_out/q_tests.txt
for sg in 60000; do
WB200_MK_STAGGER=$sg WB200_MK_TRACE=gpurun_out/q_mk_trace_$sg.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 400 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/q_bench_$sg.json 2> gpurun_out/q_bench_$sg.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/q_bench_$sg.json').read().strip().splitlines()[-1])
    print('stagger $sg:', round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('bench failed', e)
PY
tail -2 gpurun_out/q_bench_$sg.err; tail -24 gpurun_out/q_mk_trace_$sg.txt
done
for sg in 0 20000 120000; do
WB200_MK_STAGGER=$sg WB200_BENCH_REF_TOOL=0 timeout -s KILL 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/q_bench_$sg.json 2> gpurun_out/q_bench_$sg.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/q_bench_$sg.json').read().strip().splitlines()[-1])
    print('stagger $sg:', round(j['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3))
except Exception as e: print('bench failed', e)
PY
done
