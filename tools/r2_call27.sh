#!/bin/bash
mkdir -p gpurun_out
T=r2c27
timeout 300 python tools/subject_probe.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_subject_launches.csv python tools/subject_probe.py > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c27_subject_launches.csv')) if len(r)>10]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    key=(int(r[idx['ID']]), r[idx['Kernel Name']][:46], r[idx['Grid Size']])
    d.setdefault(key,{})[r[idx['Metric Name']]]=float(r[idx['Metric Value']].replace(',',''))
keys=list(d.keys())
n=len(keys)//3
last=keys[2*n:]
tot=0
for k in last:
    v=d[k]; tot+=v['gpu__time_duration.sum']
    print(k[0], k[1], k[2], round(v['gpu__time_duration.sum']/1e3,1),'us', round(v['dram__bytes_read.sum']/1e6,1),'MB r', round(v['dram__bytes_write.sum']/1e6,1),'MB w')
print('launches per rep', n, 'sum of durations (ms)', round(tot/1e6,3))
PY
