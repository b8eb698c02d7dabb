#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_train_eval.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "opnet or training_step or fused" --tb=short -x > gpurun_out/r02_61_tests.log 2>&1; tail -3 gpurun_out/r02_61_tests.log | cut -c1-200
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('early', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['kernels']['opnet_bwd_fused']['ms'])"; done
OPN_WGRAD_EARLY=0 timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('late ', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['kernels']['opnet_bwd_fused']['ms'])"
