#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k "attention or transformer" --tb=short > gpurun_out/r02_39_tests.log 2>&1; tail -4 gpurun_out/r02_39_tests.log | cut -c1-300
OPN_B200_LIB=$PWD/objectpermanence_b200/lib/libopnet_b200_phases.so timeout 120 python tools/attn_phases.py 2>&1 | tee gpurun_out/r02_39_attn_phases.log
timeout 300 python tools/attn_time.py 2>&1 | grep fused | tee gpurun_out/r02_39_attn_time.log
