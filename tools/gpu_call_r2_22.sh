#!/bin/bash
mkdir -p gpurun_out
ATTN_S=9600 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_22_attn_launches.csv python tools/attn_time.py > gpurun_out/r02_22_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_22_attn_launches.csv')) if len(r)>5]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
for r in rows[1:60]:
    n=r[idx["Kernel Name"]].split("(")[0][-60:]
    print(f'{float(r[idx["Metric Value"]])/1e3:10.1f} us  {n}')
PY
