#!/bin/bash
# round-1 late call: dropout / colred / weight-gradient-overlap checks (short: the GPU budget is nearly spent)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --tb=short -p no:cacheprovider \
    -k "dropout or attention or transformer or sgemm or skinny or golden" > gpurun_out/c16_tests.log 2>&1; tail -15 gpurun_out/c16_tests.log
echo "== skinny (new)"; timeout 60 python tools/skinny_time.py 2>&1 | tee gpurun_out/c16_skinny.log
echo "== skinny (legacy colred)"; OPN_COLRED_LEGACY=1 timeout 60 python tools/skinny_time.py 2>&1 | tee -a gpurun_out/c16_skinny.log
echo "== step A/B"; timeout 90 python tools/step_ab.py 2>&1 | tee gpurun_out/c16_step_ab.log
echo "== transformer train mode"; timeout 90 python tools/transformer_full.py 2>&1 | tee gpurun_out/c16_transformer.log
