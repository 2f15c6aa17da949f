#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2c4_suite.log 2>&1; echo "suite rc=$?"
tail -25 gpurun_out/r2c4_suite.log
timeout 300 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c4_ab.json \
  "CONV_BWD_FUSED=0" "CONV_BWD_FUSED=1" > gpurun_out/r2c4_ab.log 2>&1; echo "ab rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c4_ab.json'))
for r in d.get('interleaved',[]):
    print(r['ms_per_step_min'], r['ms_per_step_median'], r['options'])
PY
timeout 400 python bench.py --steps 500 --warmup 20 --no-library-baseline --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c4_bench.json'))
print(p['ms_per_step'], p['e2e']['ms_per_step'], p['kernels_per_step'])
for k,v in sorted(p['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']):
    print(f"{k:34s} {v['launches_per_step']:5.1f} x {v['avg_us']:7.2f} us = {v['ms_per_step']*1e3:7.1f} us")
PY
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c4_timeline.json > gpurun_out/r2c4_timeline.log 2>&1; echo "timeline rc=$?"
