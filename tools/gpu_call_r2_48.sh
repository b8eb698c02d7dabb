#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/fused_fwd_time.py 2>&1 | tail -3
OPN_OPNET_SPLIT=0 timeout 120 python tools/fused_fwd_time.py 2>&1 | tail -1 | sed 's/^/[OPN_OPNET_SPLIT=0] /'
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k "opnet or fused" --tb=short -x > gpurun_out/r02_48_tests.log 2>&1; tail -6 gpurun_out/r02_48_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
