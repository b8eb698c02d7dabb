#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py tests/test_gpu_integration.py -m gpu -q -p no:cacheprovider -k "tcgen05 or large_per_gpu or integration or no_grad_skips" > gpurun_out/r02_11_tests.log 2>&1; tail -6 gpurun_out/r02_11_tests.log
timeout 900 python bench.py > gpurun_out/r02_11_bench.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_11_bench.log').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["wall_value"])
print(json.dumps(d["readings"]["batch_sweep_one_gpu"]))
PY
