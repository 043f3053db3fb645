#!/bin/bash
# round 2: two-GPU sanity (in-library replicas + the torchrun bench path)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -2
WB200_BENCH_REF_TOOL=0 timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-ragged > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/n2_bench.json').read().splitlines() if l.startswith('{')][-1])
    print('N=2:', round(j['value'],1), round(j['e2e']['value'],1), 'n_gpus', j['n_gpus'], 'pass ms', round(j['engine']['decode_gpu_ms_per_pass'],3))
except Exception as e: print('N=2 bench failed', e)
PY
tail -2 gpurun_out/n2_bench.err
