#!/bin/bash
# Round 2, call 1: hunt the [11,37] parity failure (data vs order dependence), sanitizers, and the starting bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_01_smi.log 2>&1
HUNT_SEEDS=40 timeout 900 python tools/flake_hunt.py > gpurun_out/r02_01_hunt.log 2>&1; tail -25 gpurun_out/r02_01_hunt.log
SEQ='tests/test_gpu_kernels.py::test_opnet_fused_forward_matches_separate_kernels'
OPN_TEST_WGRAD_MODES=1 timeout 300 python -m pytest "$SEQ" -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_01_seq.log 2>&1; tail -5 gpurun_out/r02_01_seq.log
OPN_TEST_WGRAD_MODES=1 timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_01_memcheck.log \
    python -m pytest "$SEQ" -m gpu -q -x -p no:cacheprovider -k "11-37" > gpurun_out/r02_01_memcheck_run.log 2>&1; tail -8 gpurun_out/r02_01_memcheck.log
OPN_TEST_WGRAD_MODES=1 timeout 600 compute-sanitizer --tool initcheck --log-file gpurun_out/r02_01_initcheck.log \
    python -m pytest "$SEQ" -m gpu -q -x -p no:cacheprovider -k "11-37" > gpurun_out/r02_01_initcheck_run.log 2>&1; tail -8 gpurun_out/r02_01_initcheck.log
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_01_racecheck.log \
    python -m pytest "$SEQ" -m gpu -q -x -p no:cacheprovider -k "3-2" > gpurun_out/r02_01_racecheck_run.log 2>&1; tail -8 gpurun_out/r02_01_racecheck.log
timeout 300 python bench.py > gpurun_out/r02_01_bench.log 2>&1; tail -2 gpurun_out/r02_01_bench.log
