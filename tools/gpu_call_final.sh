#!/bin/bash
# end-of-round check: the whole GPU parity suite, then the bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=20 > gpurun_out/c17_tests.log 2>&1; tail -25 gpurun_out/c17_tests.log
timeout 200 python bench.py > gpurun_out/c17_bench.log 2>&1; tail -1 gpurun_out/c17_bench.log | cut -c1-2500
