#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/r02_40_suite.log 2>&1; tail -4 gpurun_out/r02_40_suite.log | cut -c1-300
python __graft_entry__.py --smoke 2>&1 | tail -2
bash tools/gpu_call_evidence.sh r02g
