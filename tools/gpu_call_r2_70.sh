#!/bin/bash
# ncu --set full of the short-K projection kernel and of head_loss (one launch each) for profiles/
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"proj_kernel" -s 3 -c 1 -o gpurun_out/r02n_proj_full -f python tools/xproj_time.py > gpurun_out/r02n_proj_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02n_proj_ncu.log | cut -c1-200
ls -la gpurun_out/r02n_proj_full.ncu-rep
