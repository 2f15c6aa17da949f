#!/bin/bash
# Round 2, GPU call 2: scheduling combinations on top of the kernels that were green and faster in call 1; fused head parity.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_trainer.py tests/test_gpu_epoch.py -q > gpurun_out/r2c2_suite.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/r2c2_suite.log
BASE="GRU_FWD_V2=1,GRU_BWD_RING=8,TN_BATCH=1,TN_STAGES=1"
timeout 500 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c2_combos.json \
  "HEAD_FUSED=0" "HEAD_FUSED=1" "$BASE" "$BASE,GRU_BWD_EXCLUSIVE_KB=200" "$BASE,WGRAD_DEFER=1" "$BASE,GRU_BWD_EXCLUSIVE_KB=200,WGRAD_DEFER=1" \
  "$BASE,GRU_BWD_EXCLUSIVE_KB=200,TN_BATCH_CTAS=40" "$BASE,GRU_BWD_EXCLUSIVE_KB=200,TN_BATCH_CTAS=96" "$BASE,GRU_BWD_EXCLUSIVE_KB=200,NT_TRIM_STAGES=1" \
  "$BASE,GRU_BWD_EXCLUSIVE_KB=200,CONV_DGRAD_V2=1,CONV_FWD_V2=1" "$BASE,GRU_BWD_EXCLUSIVE_KB=100" \
  > gpurun_out/r2c2_combos.log 2>&1; echo "combos rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c2_combos.json'))
for r in d.get('interleaved',[]):
    print(r['ms_per_step_min'], r['ms_per_step_median'], r['options'])
PY
