#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "wgrad" --tb=short -x 2>&1 | tail -3
timeout 120 python tools/wgrad_jobs_time.py 2>&1 | tee gpurun_out/r02_36_wgrad_jobs_time.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_train_eval.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/r02_36_models.log 2>&1; tail -5 gpurun_out/r02_36_models.log | cut -c1-300
