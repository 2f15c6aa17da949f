#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c17_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2c17_suite.log
timeout 300 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c17_ab.json "DROP_FUSED=0" "DROP_FUSED=1" > gpurun_out/r2c17_ab.log 2>&1; echo "ab rc=$?"; tail -2 gpurun_out/r2c17_ab.log
timeout 300 python bench.py --workload preprocess --steps 3 > gpurun_out/r2c17_pre.json 2> gpurun_out/r2c17_pre.err; echo "pre rc=$?"; tail -3 gpurun_out/r2c17_pre.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c17_pre.json'))
print('preprocess value', p['value'], 'e2e', p['e2e']['value'], 'roof', p['roofline']['frac'])
PY
timeout 200 python tools/pre_probe.py 2>&1 | head -20
