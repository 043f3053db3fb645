#!/bin/bash
# round 2, run X: the ggml backend plugin under the reference's own CLI + library
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_zz_ggml_backend_plugin_gpu.py -q -m gpu -s -x 2>&1 | grep -v "^whisper_\|^ggml_\|^model_load" | tail -30 > gpurun_out/x_tests.txt; cat gpurun_out/x_tests.txt
