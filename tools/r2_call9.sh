#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c9_suite.log 2>&1; echo "suite rc=$?"
tail -4 gpurun_out/r2c9_suite.log
timeout 500 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c9_ab.json \
  "POOL_TM_ITERS=4" "TN_STAGES=2" "TN_STAGES=3" "WGRAD_DEFER=0" "WGRAD_DEFER=0,TN_STAGES=2" "TN_STAGES=2,TN_BATCH_CTAS=128" > gpurun_out/r2c9_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/r2c9_ab.log | tail -9
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c9_timeline.json > gpurun_out/r2c9_timeline.log 2>&1; echo "timeline rc=$?"
tail -16 gpurun_out/r2c9_timeline.log
