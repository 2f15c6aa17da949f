#!/bin/bash
# final single-GPU pass on the code as shipped: suite, smoke, default bench line, driver-style short bench, reference arm, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2i_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2i_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ) 2> gpurun_out/r2i_time.txt; echo "bench rc=$?"; grep real gpurun_out/r2i_time.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_short.json 2> gpurun_out/r2i_bench_short.err; echo "short bench rc=$?"
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2i_ref.json 2> gpurun_out/r2i_ref.err ) 2> gpurun_out/r2i_time_ref.txt; echo "ref rc=$?"; grep real gpurun_out/r2i_time_ref.txt
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2i_timeline.json > gpurun_out/r2i_timeline.log 2>&1; echo "timeline rc=$?"
python - <<'PY'
import json
for f in ('r2i_bench','r2i_bench_short'):
    p=json.load(open(f'gpurun_out/{f}.json'))
    print(f, 'value', round(p['value']), 'ms', round(p['ms_per_step'],4), 'e2e', round(p['e2e']['value']), 'roof', {k:p['roofline'].get(k) for k in ('kernel','avg_us','frac','ns_per_time_step')}, 'clocks', p['clocks'])
    for k in ('loso','preprocess'):
        v=p.get(k,{})
        print('  ', k, {a:b for a,b in v.items() if a in ('value','unit','unavailable','preprocess_s','train_windows_per_s','accuracy_mean','accuracy_delta_vs_reference','e2e')})
    lg=p.get('library_gpu_baseline',{})
    print('   library', {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else None) for k,v in lg.items() if k in ('pytorch_default','fp32_strict','tf32_allowed')}, 'cpu', p.get('cpu_baseline',{}).get('value'))
r=json.load(open('gpurun_out/r2i_ref.json'))
print('reference arm', r['value'], (r.get('loso') or {}).get('train_seconds'), (r.get('loso') or {}).get('test_accuracy'))
PY
