#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py tests/test_gpu_zz_regressions.py -m gpu -q -p no:cacheprovider -s -k "transformer or attention" > gpurun_out/r02_20_tests.log 2>&1; grep -E "transformer_lstm|passed|failed|Error" gpurun_out/r02_20_tests.log | head -20
timeout 300 python tools/attn_time.py > gpurun_out/r02_20_attn_time.log 2>&1; cat gpurun_out/r02_20_attn_time.log
