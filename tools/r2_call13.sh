#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2c13_suite.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/r2c13_suite.log
timeout 300 python tools/pre_probe.py > gpurun_out/r2c13_probe.log 2>&1; echo "probe rc=$?"; head -22 gpurun_out/r2c13_probe.log
timeout 300 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c13_ab.json "TN_AFTER_NT=0" "TN_AFTER_NT=1" "TN_AFTER_NT=1,TN_STAGES=3" > gpurun_out/r2c13_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r2c13_ab.log
timeout 300 python bench.py --workload preprocess --steps 3 > gpurun_out/r2c13_pre.json 2> gpurun_out/r2c13_pre.err; echo "pre rc=$?"
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c13_pre.json'))
print('preprocess value', p['value'], 'e2e', p['e2e']['value'], 'roof', p['roofline']['frac'])
PY
timeout 300 python bench.py --workload loso --no-cpu-baseline > gpurun_out/r2c13_loso.json 2> gpurun_out/r2c13_loso.err; echo "loso rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2c13_loso.json'))
print('loso', round(l['value'],2), 's pre', round(l['preprocess_s'],2), 'windows', l['windows_trained'], 'w/s', round(l['train_windows_per_s']), 'acc', round(l['accuracy_mean'],4))
PY
