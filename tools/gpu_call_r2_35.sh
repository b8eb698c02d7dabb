#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x > gpurun_out/r02_35_suite.log 2>&1; tail -5 gpurun_out/r02_35_suite.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_35_bench.log 2>&1; tail -1 gpurun_out/r02_35_bench.log | cut -c1-600
