#!/bin/bash
mkdir -p gpurun_out
T=r2c32
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/${T}_suite.log
( time timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_time.txt; echo "bench rc=$?"; cat gpurun_out/${T}_time.txt | tail -3
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c32_bench.json'))
print({k:p[k] for k in ('value','ms_per_step','e2e','gpu_launches') if k in p})
print('roofline', p['roofline'])
for k in ('loso','preprocess','dp','library_gpu_baseline','cpu_baseline'):
    if k in p: print(k, json.dumps(p[k])[:600])
PY
