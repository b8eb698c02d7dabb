#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_04_suite.log 2>&1; tail -15 gpurun_out/r02_04_suite.log
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest tests/test_gpu_train_eval.py -m gpu -q -p no:cacheprovider -k "adam_trains_like" 2>&1 | tail -1; done > gpurun_out/r02_04_adam_repeat.log; cat gpurun_out/r02_04_adam_repeat.log
timeout 900 python bench.py > gpurun_out/r02_04_bench.log 2>&1; tail -1 gpurun_out/r02_04_bench.log | cut -c1-6000
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_04_bench_ref.log 2>&1; tail -1 gpurun_out/r02_04_bench_ref.log | cut -c1-400
