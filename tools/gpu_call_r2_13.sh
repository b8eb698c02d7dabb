#!/bin/bash
# profiler evidence: (1) smoke() under ncu's launch-list pass must complete now (round 1: ncu_rc=9); (2) launch list of the bench
# step; (3) one --set full capture of the two persistent kernels; plus the new tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_train_eval.py -m gpu -q -p no:cacheprovider -s -k "config_3_full_shape or state_dict or status_page" > gpurun_out/r02_13_tests.log 2>&1; grep -E "S=9600|passed|failed|Error" gpurun_out/r02_13_tests.log | head
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_13_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_13_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -2 gpurun_out/r02_13_smoke_ncu.log; wc -l gpurun_out/r02_13_smoke_launches.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 13400 -c 200 --csv --log-file gpurun_out/r02_13_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-readings > gpurun_out/r02_13_bench_ncu.log 2>&1; echo "ncu bench rc=$?"; wc -l gpurun_out/r02_13_bench_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:opnet_.*_fused -s 40 -c 2 -o gpurun_out/r02_13_fused_full python bench.py --steps 2 --warmup 3 --no-readings > gpurun_out/r02_13_full_ncu.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
