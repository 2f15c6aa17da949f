#!/bin/bash
# last check of the library as committed: GPU suite, smoke, a short default-path bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/r2j_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2j_smoke.log
timeout 300 python bench.py --steps 400 --warmup 10 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/r2j_b.json 2>/dev/null; echo "bench rc=$?"
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2j_b.json'))
print('ms', round(p['ms_per_step'],5), 'value', round(p['value']), 'e2e', round(p['e2e']['value']), 'launches', p['gpu_launches'], p['clocks'])
PY
