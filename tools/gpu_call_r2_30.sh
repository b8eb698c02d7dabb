#!/bin/bash
# state check after the fused attention: full GPU suite, bench line, launch list of one transformer step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/r02_30_suite.log 2>&1; tail -5 gpurun_out/r02_30_suite.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_30_bench.log 2>&1; tail -1 gpurun_out/r02_30_bench.log | cut -c1-4000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_30_tr_launches.csv python tools/profile_step.py --model transformer_lstm --steps 1 > gpurun_out/r02_30_tr_ncu.log 2>&1
tail -2 gpurun_out/r02_30_tr_ncu.log
