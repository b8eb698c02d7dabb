#!/bin/bash
# Short-K projection kernel: parity, timing against the tcgen05 path, bench line.
TAG=r02k
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --tb=short -k "sgemm or lstm_layer or opnet_fused or baseline" > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log | cut -c1-300
OPN_GEMM_PROJ=0 timeout 120 python tools/xproj_time.py 2>&1 | tail -3
timeout 120 python tools/xproj_time.py 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-300
