#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_trainer.py -m gpu -q -x > gpurun_out/r2c19_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2c19_suite.log
for pr in 0 -1 0 -1; do
  MMS_CAPTURE_PRIORITY=$pr timeout 200 python bench.py --steps 1500 --warmup 30 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/r2c19_b_$pr.json 2>/dev/null
  python - <<PY
import json
p=json.load(open('gpurun_out/r2c19_b_$pr.json'))
print('priority $pr: ms', round(p['ms_per_step'],5), 'e2e', round(p['e2e']['ms_per_step'],5))
PY
done
MMS_CAPTURE_PRIORITY=-1 timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c19_timeline.json > gpurun_out/r2c19_timeline.log 2>&1; echo "timeline rc=$?"
tail -14 gpurun_out/r2c19_timeline.log
