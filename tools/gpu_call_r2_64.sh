#!/bin/bash
# Why did smoke() under ncu stall in the previous call?  Bounded probes, each with its own time-out.
TAG=r02j
mkdir -p gpurun_out
timeout 150 python tools/smoke_traced.py > gpurun_out/${TAG}_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/${TAG}_plain.log | cut -c1-250
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_smoke_launches.csv python tools/smoke_traced.py > gpurun_out/${TAG}_ncu.log 2>&1; rc=$?; echo "ncu smoke rc=$rc"; tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-250
if [ $rc -ne 0 ]; then
  OPN_OPNET_SPLIT=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_smoke_launches_nosplit.csv python tools/smoke_traced.py > gpurun_out/${TAG}_ncu_nosplit.log 2>&1; echo "ncu nosplit rc=$?"; tail -3 gpurun_out/${TAG}_ncu_nosplit.log | cut -c1-250
else
  timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_smoke_launches_main.csv python __graft_entry__.py --smoke > gpurun_out/${TAG}_ncu_main.log 2>&1; echo "ncu __main__ rc=$?"; tail -2 gpurun_out/${TAG}_ncu_main.log | cut -c1-250
fi
