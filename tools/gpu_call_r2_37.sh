#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_zz_regressions.py -m gpu -q -p no:cacheprovider -k "attention or transformer or linear or mlp or non_linear" --tb=short > gpurun_out/r02_37_tests.log 2>&1; tail -5 gpurun_out/r02_37_tests.log | cut -c1-300
timeout 300 python tools/config3_time.py 2>&1 | tee gpurun_out/r02_37_config3_time.log
OPN_WGRAD=sgemm timeout 300 python tools/config3_time.py 2>&1 | sed 's/^/[OPN_WGRAD=sgemm] /' | tee -a gpurun_out/r02_37_config3_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_37_tr_launches.csv python tools/profile_step.py --model transformer_lstm --steps 1 > gpurun_out/r02_37_tr_ncu.log 2>&1
tail -1 gpurun_out/r02_37_tr_ncu.log
