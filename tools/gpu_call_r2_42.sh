#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_eval.py tests/test_gpu_data_parallel.py -m gpu -q -p no:cacheprovider --tb=short -x > gpurun_out/r02_42_tests.log 2>&1; tail -15 gpurun_out/r02_42_tests.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_42_bench.log 2>&1; tail -1 gpurun_out/r02_42_bench.log | cut -c1-500
OPN_HEAD_LOSS=0 timeout 900 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
