#!/bin/bash
mkdir -p gpurun_out
OPN_B200_LIB=$PWD/objectpermanence_b200/lib/libopnet_b200_phases.so timeout 120 python tools/attn_phases.py 2>&1 | tee gpurun_out/r02_38_attn_phases.log
timeout 300 python tools/attn_time.py 2>&1 | grep fused | tee gpurun_out/r02_38_attn_time.log
timeout 300 python tools/config3_time.py 2>&1 | tee gpurun_out/r02_38_config3_time.log
