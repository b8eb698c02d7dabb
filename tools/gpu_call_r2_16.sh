#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_16_tr_launches.csv python tools/profile_step.py --model transformer_lstm --steps 1 > gpurun_out/r02_16_tr_ncu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02_16_tr_ncu.log; wc -l gpurun_out/r02_16_tr_launches.csv
