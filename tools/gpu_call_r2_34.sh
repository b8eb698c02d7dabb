#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "wgrad" --tb=short -x 2>&1 | tail -3
timeout 300 python tools/wgrad_time.py 2>&1 | grep "B=" | tee gpurun_out/r02_34_wgrad_time.log
OPN_B200_LIB=$PWD/objectpermanence_b200/lib/libopnet_b200_phases.so timeout 120 python tools/wgrad_phases.py 2>&1 | tee gpurun_out/r02_34_wgrad_phases.log
