#!/bin/bash
# ncu --set full of (a) the tile/cluster backward kernels (MMS_CONV_BWD_FUSED=1) and (b) the default step's kernels
mkdir -p gpurun_out
python tools/prof_step.py 3 > gpurun_out/r2c6_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2c6_plain.log; exit 1; }
MMS_CONV_BWD_FUSED=1 timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'pool_relu_bwd_tile|conv2_dgrad|conv1_wgrad_dgate' -s 4 -c 4 \
    -o gpurun_out/r2c6_fusedbwd -f python tools/prof_step.py 3 > gpurun_out/r2c6_ncu_a.log 2>&1
echo "a rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'gru_fwd_v2|gru_bwd_ring|pool_relu_bwd_kernel|conv1d_dgrad|conv1d_wgrad|tc_gemm_tn_batch|tc_gemm_nt|attn_conv1|bn_pool_conv2|bn_relu_pool_fwd' -s 16 -c 16 \
    -o gpurun_out/r2c6_default -f python tools/prof_step.py 3 > gpurun_out/r2c6_ncu_b.log 2>&1
echo "b rc=$?"
ls -la gpurun_out/*.ncu-rep
