#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -s -k "precision" > gpurun_out/r02_12_tests.log 2>&1; grep -E "1e-2 mode|passed|failed|Error" gpurun_out/r02_12_tests.log | head -20
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_12_suite.log 2>&1; tail -5 gpurun_out/r02_12_suite.log
