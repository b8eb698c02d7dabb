#!/bin/bash
mkdir -p gpurun_out
ATTN_S=9600 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_26_attn_launches.csv python tools/attn_time.py > gpurun_out/r02_26_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_26_attn_launches.csv')) if len(r)>5]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[idx["Kernel Name"]].split("(")[0][-40:]
    a=agg.setdefault(n,[]); a.append(float(r[idx["Metric Value"]])/1e3)
for n,v in agg.items(): print(f"{sum(v)/len(v):10.1f} us x{len(v):3d}  {n}")
PY
timeout 300 python tools/attn_time.py 2>&1 | grep fused
