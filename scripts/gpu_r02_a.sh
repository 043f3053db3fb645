#!/bin/bash
# round 2, call A: pool tests + e2e tests + bench through whisper.h
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests/test_pool_gpu.py tests/test_e2e_gpu.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/a_pytest.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -5 gpurun_out/a_pytest.txt; tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
