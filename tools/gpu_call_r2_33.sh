#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/wgrad_time.py 2>&1 | tee gpurun_out/r02_33_wgrad_time.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad|gemm_tc|split_bf16|colred|sgemm|zero" -c 200 --csv --log-file gpurun_out/r02_33_wgrad_launches.csv python tools/wgrad_time.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_33_wgrad_launches.csv')) if len(r)>5]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
for r in rows[1:40]:
    print(f'{float(r[idx["Metric Value"]])/1e3:9.1f} us  {r[idx["Kernel Name"]].split("(")[0][-50:]}  grid {r[idx["Grid Size"]]}')
PY
