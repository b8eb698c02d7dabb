#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -p no:cacheprovider -k "attention or transformer or wgrad or any_hidden or unlisted" --tb=short > gpurun_out/r02_41_tests.log 2>&1; tail -4 gpurun_out/r02_41_tests.log | cut -c1-300
timeout 300 python tools/config3_time.py 2>&1 | tee gpurun_out/r02_41_config3_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"combine|reduce" -c 60 --csv --log-file gpurun_out/r02_41_small_launches.csv python tools/profile_step.py --model transformer_lstm --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_41_small_launches.csv')) if len(r)>5]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[idx["Kernel Name"]].split("(")[0][-30:]
    agg.setdefault(n,[]).append(float(r[idx["Metric Value"]])/1e3)
for n,v in agg.items(): print(f"{sum(v)/len(v):8.1f} us avg x{len(v)}  {n}")
PY
