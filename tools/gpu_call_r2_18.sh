#!/bin/bash
mkdir -p gpurun_out
OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so timeout 300 python tools/lstm_tc_phases.py > gpurun_out/r02_18_tc_phases.log 2>&1; grep "B=256" gpurun_out/r02_18_tc_phases.log
