#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "split" --tb=short 2>&1 | tail -3
timeout 100 python tools/split_fwd_debug.py > gpurun_out/r02_58_split_fwd.log 2>&1; head -1 gpurun_out/r02_58_split_fwd.log
timeout 100 python tools/split_bwd_debug.py > gpurun_out/r02_58_split_bwd.log 2>&1; head -2 gpurun_out/r02_58_split_bwd.log
timeout 900 python bench.py > gpurun_out/r02_58_bench.log 2>&1; tail -1 gpurun_out/r02_58_bench.log | cut -c1-300
