#!/bin/bash
# round 2, run S: ncu --set full (with source counters) of one launch of the second-generation decode kernel
mkdir -p gpurun_out
WB200_BENCH_REF_TOOL=0 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_pass2 --launch-skip 40 --launch-count 1 -o gpurun_out/s_mk2 \
   python bench.py --chunks 64 --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/s_ncu.log 2>&1
ls -la gpurun_out/s_mk2.ncu-rep; tail -3 gpurun_out/s_ncu.log
