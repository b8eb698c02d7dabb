#!/bin/bash
# compute-sanitizer memcheck + racecheck of the short-K projection kernel (all five shapes of its parity test)
mkdir -p gpurun_out
timeout 250 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k short_k_projection > gpurun_out/r02n_proj_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02n_proj_memcheck.log | tail -3
timeout 250 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "short_k_projection and 9600" > gpurun_out/r02n_proj_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02n_proj_racecheck.log | tail -3
