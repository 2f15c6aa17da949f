#!/bin/bash
mkdir -p gpurun_out
python tools/prof_step.py 3 > gpurun_out/r2c8_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2c8_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'pool_bwd_tm|pool_bwd_ncl|conv2_bwd|conv1_bwd|wgrad_reduce' -s 5 -c 5 \
    -o gpurun_out/r2c8_convbwd -f python tools/prof_step.py 3 > gpurun_out/r2c8_ncu.log 2>&1
echo "rc=$?"
ls -la gpurun_out/r2c8*
