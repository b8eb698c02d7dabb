#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "tcgen05" > gpurun_out/r02_10_tc_tests.log 2>&1; tail -4 gpurun_out/r02_10_tc_tests.log
TC_BATCHES=128,256,512 timeout 600 python tools/lstm_tc_time.py > gpurun_out/r02_10_tc_time.log 2>&1; grep tcgen05 gpurun_out/r02_10_tc_time.log
OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so timeout 300 python tools/lstm_tc_phases.py > gpurun_out/r02_10_tc_phases.log 2>&1; grep "B=256" gpurun_out/r02_10_tc_phases.log
OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so timeout 300 python tools/lstm_tc_skew.py > gpurun_out/r02_10_skew.log 2>&1; head -4 gpurun_out/r02_10_skew.log
