#!/bin/bash
# Final evidence of the round: full GPU suite, smoke, bench line, smoke under ncu, ncu launch list of one headline step.
TAG=r02n
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/${TAG}_suite.log 2>&1; tail -3 gpurun_out/${TAG}_suite.log | cut -c1-300
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_step_launches.csv python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_step_ncu.log 2>&1; echo "step ncu rc=$?"; python tools/ncu_step_summary.py gpurun_out/${TAG}_step_launches.csv | head -14
