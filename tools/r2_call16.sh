#!/bin/bash
# ncu --set full of every kernel of ONE default training step (eager, side streams off), and the launch list of bench.py
mkdir -p gpurun_out
python tools/prof_step.py 3 > gpurun_out/r2c16_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2c16_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'adam|attn|bn_|conv|dropout|gemm|gru|head|pool|transpose|wgrad' -s 66 -c 33 \
    -o gpurun_out/r2c16_step -f python tools/prof_step.py 3 > gpurun_out/r2c16_ncu.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r2c16*
