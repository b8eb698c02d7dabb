#!/bin/bash
# Projection kernel (batch 16, 37 CTAs per column tile) and the SM limit of the in-shadow weight-gradient launch: parity subset,
# timing, bench with and without the limit.
TAG=r02m
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_eval.py -m gpu -q -p no:cacheprovider --tb=short -k "sgemm or wgrad or opnet or head_loss or training_step" > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log | cut -c1-300
timeout 120 python tools/xproj_time.py 2>&1 | tail -3
OPN_WGRAD_SM_LIMIT=0 timeout 600 python bench.py --no-readings --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_nolimit.log 2>&1; tail -1 gpurun_out/${TAG}_bench_nolimit.log | cut -c1-230
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-230
OPN_WGRAD_SM_LIMIT=0 timeout 600 python bench.py --no-readings --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-230
timeout 600 python bench.py --no-readings --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-230
