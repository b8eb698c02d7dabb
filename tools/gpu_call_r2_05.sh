#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:cacheprovider -k "tcgen05" > gpurun_out/r02_05_tc_tests.log 2>&1; tail -30 gpurun_out/r02_05_tc_tests.log
timeout 600 python tools/lstm_tc_time.py > gpurun_out/r02_05_tc_time.log 2>&1; tail -30 gpurun_out/r02_05_tc_time.log
