#!/bin/bash
mkdir -p gpurun_out
OPN_B200_LIB=objectpermanence_b200/lib/libopnet_b200_phases.so timeout 300 python tools/attn_phases.py > gpurun_out/r02_23_attn_phases.log 2>&1; cat gpurun_out/r02_23_attn_phases.log
