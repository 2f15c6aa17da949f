#!/bin/bash
mkdir -p gpurun_out
T=r2c31
timeout 600 ncu --launch-skip 20 --launch-count 20 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${T}_resample_launches.csv python tools/resample_probe.py > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c31_resample_launches.csv')) if len(r)>10]
hdr=rows[0]; 
idx={h:i for i,h in enumerate(hdr)}
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    key=(r[idx['ID']], r[idx['Kernel Name']][:40], r[idx['Grid Size']] if 'Grid Size' in idx else '')
    d.setdefault(key,{})[r[idx['Metric Name']]]=(r[idx['Metric Value']], r[idx['Metric Unit']])
for k,v in d.items():
    print(k, {m:v[m] for m in v})
PY
