#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "fused_attention" --tb=line > gpurun_out/r02_21_attn_tests.log 2>&1; tail -16 gpurun_out/r02_21_attn_tests.log | cut -c1-250
timeout 300 python tools/attn_time.py > gpurun_out/r02_21_attn_time.log 2>&1; cat gpurun_out/r02_21_attn_time.log
