#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_trainer.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2 3; do timeout 300 python bench.py --workload preprocess --steps $((i==3 ? 20 : 400)) > gpurun_out/r2c62_pre_$i.json 2>/dev/null; python - <<PY
import json
p=json.load(open('gpurun_out/r2c62_pre_$i.json'))
print('preprocess', round(p['value'],1), 'subjects/s  e2e', round(p['e2e']['value'],1), 'frac', round(p['roofline']['frac'],4), 'reps', p['steps'])
PY
done
