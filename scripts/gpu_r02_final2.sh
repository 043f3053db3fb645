#!/bin/bash
# round 2, final evidence: ncu --set full of one launch of k_decode_pass (64 rows) + the launch list of one whole default bench step
mkdir -p gpurun_out
WB200_BENCH_REF_TOOL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'^k_decode_pass$' --launch-skip 30 --launch-count 1 -o gpurun_out/final_mk \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/final_ncu_mk.log 2>&1
ls -la gpurun_out/final_mk.ncu-rep
WB200_BENCH_REF_TOOL=0 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final_launches.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/final_ncu_launches.log 2>&1
wc -l gpurun_out/final_launches.csv
