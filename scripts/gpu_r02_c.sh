#!/bin/bash
# round 2, call C: phase trace (with GEMV sub-phases) of the persistent decode kernel + tensor-pipe metrics of the encoder kernels
mkdir -p gpurun_out
WB200_MK_TRACE=gpurun_out/c_mk_trace.txt timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
cat gpurun_out/c_mk_trace.txt | tail -30
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none -k regex:"gemm_kernel|fattn" -c 460 --csv --log-file gpurun_out/c_ncu_encoder.csv python scripts/enc_one.py > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log; wc -l gpurun_out/c_ncu_encoder.csv
