#!/bin/bash
# One gpurun call that (re)validates everything on a B200 box and leaves the evidence in gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_validate.sh'
# (each gpurun call costs ~30 s of box time on top of the command: batch checks instead of calling per test)
mkdir -p gpurun_out
# 1. the GPU test suite WITHOUT -x, so that one failure does not hide the rest; files that were never run on a GPU sort last
timeout -s KILL 1000 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^whisper_\|^model_load\|^vad_\|^ggml_" | tail -120 > gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
# 2. smoke + the default bench line (+ the per-phase trace of the persistent decode kernel)
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
WB200_MK_TRACE=gpurun_out/mk_trace.txt timeout -s KILL 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
