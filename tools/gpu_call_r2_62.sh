#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_data_parallel.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -3
for e in 1 0; do OPN_DP_EARLY_BUCKET=$e timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$e bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('early bucket $e:', round(d['value']), d['ms_per_step'], d.get('grads_bit_identical_across_ranks'))"; done
