#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_eval.py -m gpu -q -p no:cacheprovider --tb=short -x -k "head_loss or training_step" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_44_step_launches.csv python tools/profile_step.py --steps 2 > /dev/null 2>&1; python tools/ncu_step_summary.py gpurun_out/r02_44_step_launches.csv | head -4
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
