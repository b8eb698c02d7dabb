#!/bin/bash
mkdir -p gpurun_out
HUNT_SEEDS=150 timeout 1200 python tools/flake_hunt2.py > gpurun_out/r02_02_hunt2.log 2>&1; tail -12 gpurun_out/r02_02_hunt2.log
OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so timeout 300 python tools/lstm_phases.py > gpurun_out/r02_02_phases.log 2>&1; tail -12 gpurun_out/r02_02_phases.log
