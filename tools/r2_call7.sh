#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c7_suite.log 2>&1; echo "suite rc=$?"
tail -15 gpurun_out/r2c7_suite.log
timeout 300 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c7_ab.json \
  "CONV_BWD=0" "CONV_BWD=1" "CONV_BWD=1,CONV1_BWD_CH=960" "CONV_BWD=1,CONV1_BWD_CH=256" > gpurun_out/r2c7_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/r2c7_ab.log | tail -8
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c7_timeline.json > gpurun_out/r2c7_timeline.log 2>&1; echo "timeline rc=$?"
cat gpurun_out/r2c7_timeline.log
