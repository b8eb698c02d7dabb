#!/bin/bash
TAG=r02h
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_step_launches.csv python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_step_ncu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${TAG}_step_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_config3_launches.csv python tools/profile_step.py --model transformer_lstm --steps 1 > gpurun_out/${TAG}_config3_ncu.log 2>&1; echo "rc=$?"
timeout 300 python tools/config3_time.py 2>&1 | tee gpurun_out/${TAG}_config3_time.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
