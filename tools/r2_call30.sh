#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/preprocess_order_probe.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --workload preprocess > gpurun_out/r2c30_preprocess.json 2>gpurun_out/r2c30_preprocess.err; echo "bench preprocess rc=$?"; python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c30_preprocess.json'))
print({k:p[k] for k in ('value','unit','e2e') if k in p}); print(p.get('roofline')['frac'])
PY
