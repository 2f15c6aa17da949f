#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_trainer.py tests/test_gpu_epoch.py -m gpu -q -x > gpurun_out/r2c14_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2c14_suite.log
timeout 300 python tools/pre_probe.py > gpurun_out/r2c14_probe.log 2>&1; echo "probe rc=$?"; head -17 gpurun_out/r2c14_probe.log
timeout 300 python bench.py --workload preprocess --steps 3 > gpurun_out/r2c14_pre.json 2> gpurun_out/r2c14_pre.err; echo "pre rc=$?"; tail -3 gpurun_out/r2c14_pre.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c14_pre.json'))
print('preprocess value', p['value'], 'e2e', p['e2e']['value'], 'roof', p['roofline']['frac'])
PY
timeout 300 python bench.py --workload loso --no-cpu-baseline > gpurun_out/r2c14_loso.json 2> gpurun_out/r2c14_loso.err; echo "loso rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2c14_loso.json'))
print('loso', round(l['value'],2), 's pre', round(l['preprocess_s'],2), 'windows', l['windows_trained'], 'w/s', round(l['train_windows_per_s']), 'acc', round(l['accuracy_mean'],4))
PY
