#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2c12_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2c12_suite.log
timeout 300 python tools/pre_probe.py > gpurun_out/r2c12_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2c12_probe.log
timeout 200 python tools/graph_timeline.py --out gpurun_out/r2c12_timeline.json > gpurun_out/r2c12_timeline.log 2>&1; echo "timeline rc=$?"
cat gpurun_out/r2c12_timeline.log
