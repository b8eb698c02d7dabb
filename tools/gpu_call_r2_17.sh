#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "tcgen05" > gpurun_out/r02_17_tc_tests.log 2>&1; tail -3 gpurun_out/r02_17_tc_tests.log
echo "== RED (L2 vector atomics)"; TC_BATCHES=256,512 timeout 600 python tools/lstm_tc_time.py 2>&1 | grep "tcgen05 split"
echo "== stored blocks"; OPN_LSTM_TC_RED=0 TC_BATCHES=256,512 timeout 600 python tools/lstm_tc_time.py 2>&1 | grep "tcgen05 split"
