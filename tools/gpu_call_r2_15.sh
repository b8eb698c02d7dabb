#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_data_parallel.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_15_dp_test.log 2>&1; tail -3 gpurun_out/r02_15_dp_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_15_bench2.log 2>&1; tail -1 gpurun_out/r02_15_bench2.log | cut -c1-1500
