#!/bin/bash
mkdir -p gpurun_out
T=r2c55
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/${T}_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/${T}_suite.log
timeout 200 python tools/graph_timeline.py --out gpurun_out/${T}_timeline.json > gpurun_out/${T}_timeline.log 2>&1; echo "timeline rc=$?"
grep -E "span|tc_gemm" gpurun_out/${T}_timeline.log
for i in 1 2; do timeout 200 python bench.py --steps 1500 --warmup 30 --no-subrecords --no-cpu-baseline --no-library-baseline > gpurun_out/${T}_b.json 2>/dev/null
python - <<PY
import json
p=json.load(open('gpurun_out/${T}_b.json'))
print('ms', round(p['ms_per_step'],5), 'e2e', round(p['e2e']['ms_per_step'],5), {k:round(v['avg_us'],1) for k,v in p['kernels'].items() if 'tc_gemm' in k})
PY
done
