#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --workload loso --no-cpu-baseline --concurrent-folds $k > gpurun_out/r2c11_loso_$tag.json 2> gpurun_out/r2c11_loso_$tag.err; python - <<PY
import json
l=json.load(open('gpurun_out/r2c11_loso_$tag.json'))
print('$tag', round(l['value'],2), 's pre', round(l['preprocess_s'],2), 'windows', l['windows_trained'], 'w/s', round(l['train_windows_per_s']), 'acc', round(l['accuracy_mean'],4))
PY
}
run default 4 MMS_NOP=1
run tn1 4 MMS_TN_STAGES=1
run convbwd0 4 MMS_CONV_BWD=0
run convbwd0_tn1 4 MMS_CONV_BWD=0 MMS_TN_STAGES=1
run k6 6 MMS_NOP=1
run k3 3 MMS_NOP=1
run k1 1 MMS_NOP=1
