#!/bin/bash
# 8 GPUs: the data-parallel bench line (NCCL, early bucket), once
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02n_bench_${N}gpu.log 2>&1; tail -1 gpurun_out/r02n_bench_${N}gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N:', round(d['value']), d['ms_per_step'], d.get('grads_bit_identical_across_ranks'), round(d['e2e']['value']))"
