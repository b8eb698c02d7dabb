#!/bin/bash
mkdir -p gpurun_out
# launch list of the headline step: all launches of a short run (5 warm-up + 2 steps); the summary script takes the last step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_14_step_launches.csv python tools/profile_step.py --steps 2 > gpurun_out/r02_14_step_ncu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02_14_step_ncu.log; wc -l gpurun_out/r02_14_step_launches.csv
# the batch-wide tcgen05 kernels: full capture, one launch each
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_.*_tc_kernel -s 2 -c 2 -o gpurun_out/r02_14_lstm_tc_full python tools/profile_step.py --tc --steps 1 > gpurun_out/r02_14_tc_ncu.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r02_14_lstm_tc_full.ncu-rep
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_14_bench.log 2>&1; tail -1 gpurun_out/r02_14_bench.log | cut -c1-700
