#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c20_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2c20_suite.log
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c20_timeline.json > gpurun_out/r2c20_timeline.log 2>&1; echo "timeline rc=$?"
head -8 gpurun_out/r2c20_timeline.log
timeout 200 python bench.py --steps 1500 --warmup 30 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/r2c20_b.json 2>/dev/null
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c20_b.json'))
print('ms', round(p['ms_per_step'],5), 'e2e', round(p['e2e']['ms_per_step'],5))
PY
