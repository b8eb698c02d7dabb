#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/r02_45_suite.log 2>&1; tail -4 gpurun_out/r02_45_suite.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
