#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dbg_beam.py 1 16 0.001 1 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | tail -14 > gpurun_out/g_dbg16.txt; cat gpurun_out/g_dbg16.txt
timeout 300 python scripts/dbg_beam.py 6 0 0.0 0 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | tail -14 > gpurun_out/g_dbgq.txt; cat gpurun_out/g_dbgq.txt
WB200_MK_TRACE=gpurun_out/g_mk_trace.txt WB200_BENCH_REF_TOOL=0 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
tail -22 gpurun_out/g_mk_trace.txt | head -14
python - <<'PY'
import json
j=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], j['engine']['decode_gpu_ms_per_pass'], j['roofline']['frac'], j['encode_ms'], j['engine']['encode_gpu_ms_per_window'], j['encode_roofline'], j['ragged'])
PY
tail -3 gpurun_out/g_bench.err
