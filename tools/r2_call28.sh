#!/bin/bash
mkdir -p gpurun_out
T=r2c28
timeout 600 python -m pytest tests/test_gpu_preprocess.py -m gpu -q -x > gpurun_out/${T}_pp.log 2>&1; echo "preprocess tests rc=$?"; tail -5 gpurun_out/${T}_pp.log
for fast in 1 0; do
MMS_RESAMPLE_FAST=$fast timeout 300 python - <<'PY'
import os, torch, time
from multimodalsignal_b200 import preprocess as pp, _ext
lib = _ext.lib()
n = 4200000 + 137*7
num = pp.resampled_length(n, 700, 64)
x = torch.randn(8, n, dtype=torch.float64, device="cuda")
for _ in range(2): y = pp.resample_on_device(x, num)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y = pp.resample_on_device(x, num)
e1.record(); torch.cuda.synchronize()
print("FAST", os.environ.get("MMS_RESAMPLE_FAST"), "chest 8 x", n, "->", num, ":", round(e0.elapsed_time(e1)/5, 3), "ms")
import ctypes as C
lib.mms_profile_enable(1)
y = pp.resample_on_device(x, num)
torch.cuda.synchronize()
buf = (C.c_char * 16384)()
lib.mms_profile_report(buf, 16384)
lib.mms_profile_enable(0)
for line in buf.value.decode().strip().splitlines():
    name, cnt, tot = line.rsplit(" ", 2)
    print("   ", name, cnt, round(1e3*float(tot),1), "us total", round(1e3*float(tot)/int(cnt),1), "us each")
PY
done
timeout 300 python bench.py --workload preprocess > gpurun_out/${T}_preprocess.json 2>gpurun_out/${T}_preprocess.err; echo "bench preprocess rc=$?"; python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c28_preprocess.json'))
print({k:p[k] for k in ('value','unit','e2e') if k in p}); print(p.get('roofline'))
PY
timeout 300 python tools/subject_probe.py 2>&1 | tail -3
