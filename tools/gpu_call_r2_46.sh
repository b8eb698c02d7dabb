#!/bin/bash
# 2 GPUs: the NCCL gradient-equivalence test and the data-parallel bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_data_parallel.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02n_bench_2gpu.log 2>&1; tail -1 gpurun_out/r02n_bench_2gpu.log | cut -c1-1500
