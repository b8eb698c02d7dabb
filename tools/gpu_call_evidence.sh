#!/bin/bash
# evidence pass: launch lists of the headline and config-3 steps, full ncu captures of the new kernels, the bench line
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_step_launches.csv python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_step_ncu.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${TAG}_step_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_config3_launches.csv python tools/profile_step.py --model transformer_lstm --steps 1 > gpurun_out/${TAG}_config3_ncu.log 2>&1; echo "rc=$?"
ATTN_ONLY=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)" -s 12 -c 3 -o gpurun_out/${TAG}_attention_full python tools/attn_time.py > gpurun_out/${TAG}_attn_ncu.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 2 -c 1 -o gpurun_out/${TAG}_wgrad_full python tools/wgrad_time.py > gpurun_out/${TAG}_wgrad_ncu.log 2>&1; echo "rc=$?"
ls -la gpurun_out/${TAG}_*.ncu-rep
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-400
