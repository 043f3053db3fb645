#!/bin/bash
# round 2, call I: the whole GPU suite, then profiler captures for profiles/
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^whisper_\|^model_load\|^vad_\|^ggml_" | tail -60 > gpurun_out/i_gpu_tests.log
tail -6 gpurun_out/i_gpu_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/i_smoke.log 2>&1; tail -2 gpurun_out/i_smoke.log
# encoder kernels of a 16-window batched encode: durations + tensor-pipe activity
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none -k regex:"gemm2_kernel|gemm_kernel|fattn|layernorm|dequant|mel_window" -c 700 --csv --log-file gpurun_out/i_ncu_encoder16.csv \
   python bench.py --chunks 16 --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/i_ncu_enc.log 2>&1
wc -l gpurun_out/i_ncu_encoder16.csv
# one 64-row launch of the persistent decode kernel, full set
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_pass --launch-skip 30 --launch-count 1 -o gpurun_out/i_mk_r02 \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ragged > gpurun_out/i_ncu_mk.log 2>&1
ls -la gpurun_out/i_mk_r02.ncu-rep
