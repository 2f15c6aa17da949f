#!/bin/bash
mkdir -p gpurun_out
T=r2c48
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -q -x > gpurun_out/${T}_pp.log 2>&1; echo "preprocess tests rc=$?"; tail -2 gpurun_out/${T}_pp.log
for fs in 1 0; do
MMS_RESAMPLE_FILTER_STREAM=$fs timeout 300 python - <<'PY'
import os, torch
from multimodalsignal_b200 import preprocess as pp
n = 4200000 + 137*7
num = pp.resampled_length(n, 700, 64)
x = torch.randn(8, n, dtype=torch.float64, device="cuda")
for _ in range(3): y = pp.resample_on_device(x, num)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = pp.resample_on_device(x, num)
e1.record(); torch.cuda.synchronize()
print("FILTER_STREAM", os.environ.get("MMS_RESAMPLE_FILTER_STREAM"), "chest:", round(e0.elapsed_time(e1)/10, 3), "ms")
PY
done
timeout 600 python tools/preprocess_order_probe.py 2>&1 | tail -9
timeout 300 python bench.py --workload preprocess > gpurun_out/${T}_preprocess.json 2>gpurun_out/${T}_preprocess.err; echo "bench preprocess rc=$?"; python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c48_preprocess.json'))
print({k:p[k] for k in ('value','unit','e2e') if k in p}); print(p.get('roofline')['frac'])
PY
