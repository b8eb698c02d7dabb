#!/bin/bash
mkdir -p gpurun_out
echo "== diag (default)"; timeout 60 python tools/grad_flake_diag.py 2>&1 | tail -40 | cut -c1-600 | tee gpurun_out/c18_diag.log
echo "== the two tests again"; timeout 60 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_eval.py -m gpu -q --tb=line -p no:cacheprovider -k "fused_forward_matches or fused_adam_trains" 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/c18_tests.log
