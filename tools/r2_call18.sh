#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c18_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2c18_suite.log
timeout 300 python tools/ab_variants.py --no-parity --interleave 3 --steps 300 --out gpurun_out/r2c18_ab.json "NT_STAGES=2" "NT_STAGES=4" "NT_STAGES=4,DROP_FUSED=0" > gpurun_out/r2c18_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r2c18_ab.log
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c18_timeline.json > gpurun_out/r2c18_timeline.log 2>&1; echo "timeline rc=$?"
cat gpurun_out/r2c18_timeline.log
