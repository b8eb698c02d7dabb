#!/bin/bash
# First GPU call of round 2: root-cause the [B=11, T=37] in-line parity failure recorded in DESIGN.md section 9.
#   1. the exact failing sequence of the test-suite (the two-stream case, then the in-line case, same process)
#   2. the same under compute-sanitizer memcheck and initcheck (out-of-bounds / uninitialised global reads); the polling
#      time-outs of the persistent kernels are in clock cycles, so a sanitizer slow-down does not trip them early
mkdir -p gpurun_out
SEQ='tests/test_gpu_kernels.py::test_opnet_fused_forward_matches_separate_kernels'
export OPN_TEST_WGRAD_MODES=1      # tests/test_gpu_kernels.py adds the "fused_overlap" + "fused_inline_after" cases when set
timeout 300 python -m pytest "$SEQ" -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r02_seq.log 2>&1; tail -5 gpurun_out/r02_seq.log
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log \
    python -m pytest "$SEQ" -m gpu -q -x -p no:cacheprovider -k "11-37" > gpurun_out/r02_memcheck_run.log 2>&1; tail -20 gpurun_out/r02_memcheck.log
timeout 900 compute-sanitizer --tool initcheck --log-file gpurun_out/r02_initcheck.log \
    python -m pytest "$SEQ" -m gpu -q -x -p no:cacheprovider -k "11-37" > gpurun_out/r02_initcheck_run.log 2>&1; tail -20 gpurun_out/r02_initcheck.log
