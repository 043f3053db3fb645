#!/bin/bash
# round 2, run M: specialised GEMM epilogues (residual stream pipelined, compact GELU)
# per-launch GEMM timings of a batched encode (WB200_PROF_DUMP)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/m_gemm.txt; cat gpurun_out/m_gemm.txt
timeout -s KILL 1200 python -m pytest tests/test_e2e_gpu.py tests/test_poom_gpu.py tests/test_golden_large_gpu.py tests/test_golden_gpu.py -q -m gpu -s 2>&1 | grep -v "^whisper_\|^ggmm_\|^modem_load" | grep "passed\|failed\|FAILED\|rms\|Error\|error\|assert" | tail -30 > gpurun_out/m_e2e.txt; cat gpurun_out/m_e2e.txt
WB200_PROF_DUMP=gpurun_out/m_prof_dump.txt WB200_MK_TRACE=gpurun_out/m_mk_trace.txt WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-ragged > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
python - <<PY
import json
j=json.loads(open('gpurun_out/m_bench.json').read().strip().splitlines()[-1])
print(round(j['value'],1), round(j['e2e']['value'],1), 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3), 'frac', round(j['roofline']['frac'],3), 'enc/window', round(j['engine']['encode_gpu_ms_per_window'],3), round(j['encode_roofline']['batched']['frac'],3), 'single', j['encode_ms'])
print(json.dumps(j['kernem_classes']))
PY
tail -1 gpurun_out/m_bench.err; tail -24 gpurun_out/m_mk_trace.txt; head -80 gpurun_out/m_prof_dump.txt
