#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 1000 --warmup 20 > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err ) 2> gpurun_out/r2c10_time.txt; echo "bench rc=$?"; cat gpurun_out/r2c10_time.txt
tail -5 gpurun_out/r2c10_bench.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c10_bench.json'))
print('ms', p['ms_per_step'], 'e2e', p['e2e']['ms_per_step'], 'roof', {k:p['roofline'].get(k) for k in ('kernel','avg_us','frac','traffic','ns_per_time_step')})
for k in ('loso','preprocess','dp'):
    v=p.get(k,{})
    print(k, {a:b for a,b in v.items() if a in ('value','unit','unavailable','preprocess_s','train_windows_per_s','accuracy_mean','accuracy_delta_vs_reference','e2e')})
fv=p.get('loso',{}).get('fold_vs_reference',{})
print(json.dumps(fv)[:1500])
print(p.get('library_gpu_baseline',{}).get('fp32_strict'), p.get('cpu_baseline',{}).get('value'))
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2c10_ref.json 2> gpurun_out/r2c10_ref.err ) 2> gpurun_out/r2c10_time_ref.txt; echo "ref rc=$?"; cat gpurun_out/r2c10_time_ref.txt
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c10_ref.json'))
print(p['value'], p['cpu_baseline']['cores'], json.dumps(p.get('loso'))[:800])
PY
