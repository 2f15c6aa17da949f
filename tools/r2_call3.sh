#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2c3_suite.log 2>&1; echo "suite rc=$?"
tail -15 gpurun_out/r2c3_suite.log
timeout 300 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c3_ab.json \
  "CONV_FUSED=0" "CONV_FUSED=1" > gpurun_out/r2c3_ab.log 2>&1; echo "ab rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3_ab.json'))
for r in d.get('interleaved',[]):
    print(r['ms_per_step_min'], r['ms_per_step_median'], r['options'])
PY
tail -5 gpurun_out/r2c3_ab.log
