#!/bin/bash
mkdir -p gpurun_out
python tools/prof_step.py 3 > gpurun_out/r2c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'attn_conv1|bn_pool_conv2|pool_relu_bwd_tile|conv2_dgrad|conv1_wgrad_dgate|tc_gemm_tn_batch' -s 8 -c 8 \
    -o gpurun_out/r2c5_conv -f python tools/prof_step.py 3 > gpurun_out/r2c5_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2c5_plain.log; tail -5 gpurun_out/r2c5_ncu.log
