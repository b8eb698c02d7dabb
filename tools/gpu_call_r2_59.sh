#!/bin/bash
mkdir -p gpurun_out
for c in "40 50" "128 300" "32 300"; do set -- $c; BB=$1 TT=$2 timeout 100 python tools/split_fwd_debug.py 2>&1 | grep -v "^  [a-z]" | cut -c1-200; BB=$1 TT=$2 timeout 100 python tools/split_fwd_debug.py 2>&1 | grep "e-0[0-5]\|NaNs [1-9]" | head -3; BB=$1 TT=$2 timeout 100 python tools/split_bwd_debug.py 2>&1 | head -1 | cut -c1-200; BB=$1 TT=$2 timeout 100 python tools/split_bwd_debug.py 2>&1 | grep "dgates1\|dlogits" | cut -c1-120; done
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k "opnet or fused or split" --tb=short > gpurun_out/r02_59_tests.log 2>&1; tail -3 gpurun_out/r02_59_tests.log | cut -c1-200
