#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "fused_attention" --tb=line > gpurun_out/r02_25_attn_tests.log 2>&1; tail -6 gpurun_out/r02_25_attn_tests.log | cut -c1-250
timeout 300 python tools/attn_time.py 2>&1 | grep fused
