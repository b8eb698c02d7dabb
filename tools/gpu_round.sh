#!/bin/bash
# standard GPU round: micro-benchmarks (optional), timing of the recurrence kernels, parity tests, bench
TAG=$1
mkdir -p gpurun_out
if [ -n "$MICRO" ]; then
  echo "== dsmem exchange"; timeout 120 tools/bin/dsmem_exchange_bench 2>&1 | tee gpurun_out/${TAG}_dsmem.log
  echo "== hmma probe"; timeout 120 tools/bin/hmma_probe 2>&1 | tee gpurun_out/${TAG}_hmma.log
fi
echo "== timing (default)"; BS=32 timeout 200 python tools/lstm_time.py 2>&1 | tee gpurun_out/${TAG}_time.log
echo "== timing (ffma)"; OPN_LSTM_MATH=ffma BS=32 timeout 200 python tools/lstm_time.py 2>&1 | tee -a gpurun_out/${TAG}_time.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s --maxfail=10 --tb=short -p no:cacheprovider ${KEXPR:+-k "$KEXPR"} > gpurun_out/${TAG}_kernels.log 2>&1; tail -40 gpurun_out/${TAG}_kernels.log
if [ -z "$SKIP_MODELS" ]; then
timeout 1500 python -m pytest tests/test_gpu_models.py tests/test_gpu_data_parallel.py -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider > gpurun_out/${TAG}_models.log 2>&1; tail -6 gpurun_out/${TAG}_models.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; tail -2 gpurun_out/${TAG}_bench.log | cut -c1-1800
fi
