#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "wgrad" --tb=short -x > gpurun_out/r02_32_tests.log 2>&1; tail -25 gpurun_out/r02_32_tests.log | cut -c1-250
timeout 300 python tools/wgrad_time.py 2>&1 | tee gpurun_out/r02_32_wgrad_time.log
