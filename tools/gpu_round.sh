#!/bin/bash
# standard GPU round: timing of the recurrence kernels, parity tests, bench
TAG=$1
mkdir -p gpurun_out
echo "== timing (default knobs)"; BS=32 timeout 200 python tools/lstm_time.py 2>&1 | tee gpurun_out/${TAG}_time.log
echo "== timing (poll-all)"; OPN_LSTM_POLL_ALL=1 BS=32 timeout 200 python tools/lstm_time.py 2>&1 | tee -a gpurun_out/${TAG}_time.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_kernels.log 2>&1; tail -6 gpurun_out/${TAG}_kernels.log
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_models.log 2>&1; tail -6 gpurun_out/${TAG}_models.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; tail -2 gpurun_out/${TAG}_bench.log | cut -c1-1500
