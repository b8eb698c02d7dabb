#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_zz_regressions.py -m gpu -q -p no:cacheprovider -k "attention or transformer" --tb=short > gpurun_out/r02_31_tests.log 2>&1; tail -8 gpurun_out/r02_31_tests.log | cut -c1-250
timeout 300 python tools/attn_time.py 2>&1 | grep fused | tee gpurun_out/r02_31_attn_time.log
