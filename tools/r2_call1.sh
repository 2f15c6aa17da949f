#!/bin/bash
# Round 2, GPU call 1: full GPU suite (new whole-step parity tests), bench line with the stock PyTorch/cuDNN baseline,
# the round-2 sweep of every switch left unmeasured at the end of round 1.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c1_suite.log 2>&1; echo "suite rc=$?" 
tail -5 gpurun_out/r2c1_suite.log
timeout 600 python bench.py --steps 1000 --warmup 20 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/r2c1_bench.json
timeout 700 python tools/round2_sweep.py > gpurun_out/r2c1_sweep.log 2>&1; echo "sweep rc=$?"
tail -60 gpurun_out/r2c1_sweep.log
