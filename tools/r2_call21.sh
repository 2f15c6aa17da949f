#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/r2c21_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2c21_suite.log
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c21_timeline.json > gpurun_out/r2c21_timeline.log 2>&1; echo "timeline rc=$?"
grep -E "span|gru_" gpurun_out/r2c21_timeline.log
for i in 1 2; do timeout 200 python bench.py --steps 1500 --warmup 30 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/r2c21_b.json 2>/dev/null
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c21_b.json'))
print('ms', round(p['ms_per_step'],5), 'e2e', round(p['e2e']['ms_per_step'],5), {k:round(v['avg_us'],1) for k,v in p['kernels'].items() if k.startswith('gru')})
PY
done
