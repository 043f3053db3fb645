#!/bin/bash
# round 2, final validation: whole GPU suite, smoke, default bench line with phase trace, ncu --set full of the decode kernel, launch list
mkdir -p gpurun_out
bash scripts/gpu_validate.sh
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_pass<" --launch-skip 30 --launch-count 1 -o gpurun_out/final_mk \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/final_ncu_mk.log 2>&1
ls -la gpurun_out/final_mk.ncu-rep
WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --chunks 16 --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/final_ncu_launches.log 2>&1
wc -l gpurun_out/final_launches.csv
