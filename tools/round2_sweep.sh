#!/bin/bash
# First GPU call of round 2: everything the end of round 1 left unmeasured, ~3 minutes on one B200.
#   gpurun --timeout 420 -- 'bash tools/round2_sweep.sh'
# 1. parity + eager kernel times of the switches that select different KERNELS (each variant: GRU op cases, golden model
#    cases, smoke() at the tcgen05-sized batch; a variant whose parity fails is not timed)
# 2. the batched TN kernel's own op test (written without GPU access)
# 3. interleaved graph timings of the side-stream experiments against the default
set -u
mkdir -p gpurun_out
timeout 100 python tools/ab_variants.py --quick --out gpurun_out/r2_parity.json \
    TN_STAGES=1 TN_SPLIT=48 WGRAD1_TILE=480 NT_TRIM_STAGES=1 GRU_BWD_RING=8 CONV_FWD_V2=1 CONV_DGRAD_V2=1 GRU_FWD_V2=1 TN_BATCH=1 TN_BATCH=1,TN_STAGES=1 \
    > gpurun_out/r2_parity.log 2>&1
MMS_TEST_EXPERIMENTAL=1 timeout 60 python -m pytest tests/test_gpu_tc_gemm.py -q -k batch > gpurun_out/r2_tn_batch_test.log 2>&1
MMS_CONV_FWD_V2=1 MMS_CONV_DGRAD_V2=1 timeout 60 python -m pytest tests/test_gpu_ops.py -q -k conv1d >> gpurun_out/r2_tn_batch_test.log 2>&1
timeout 120 python tools/ab_variants.py --interleave 3 --steps 300 --out gpurun_out/r2_interleaved.json \
    GRU_BWD_RING=4 TN_STAGES=1 TN_SPLIT=48 TN_STAGES=1,TN_SPLIT=48 WGRAD1_TILE=480 TN_STAGES=1,WGRAD1_TILE=480 \
    GRU_BWD_EXCLUSIVE_KB=200 WGRAD_DEFER=1 NT_TRIM_STAGES=1 GRU_BWD_RING=8 CONV_FWD_V2=1 CONV_DGRAD_V2=1 GRU_FWD_V2=1 GRU_FWD_V2=1,CONV_FWD_V2=1,CONV_DGRAD_V2=1,GRU_BWD_RING=8 TN_BATCH=1 TN_BATCH=1,WGRAD1_TILE=480 \
    SIDE_STREAMS=0,GRU_BWD_RING=4 \
    > gpurun_out/r2_interleaved.log 2>&1
# 3b. the programmatic-dependent-launch build (libmms_b200_pdl.so): parity, then its step time against r2_parity.json's baseline
MMS_B200_LIB=$PWD/multimodalsignal_b200/libmms_b200_pdl.so timeout 60 python tools/ab_variants.py --quick \
    --out gpurun_out/r2_pdl.json GRU_BWD_RING=4 > gpurun_out/r2_pdl.log 2>&1
# 3c. the whole GPU suite with the depth-8 backward ring (66.7 us against 69.3 us per launch at depth 4): green = it may become the default
MMS_GRU_BWD_RING=8 timeout 90 python -m pytest tests -m gpu -x -q > gpurun_out/r2_suite_ring8.log 2>&1
# 4. where the time goes INSIDE the graph: per-launch start / end with the side streams on and off (stretch per kernel)
timeout 60 python tools/graph_timeline.py --out gpurun_out/r2_timeline.json > gpurun_out/r2_timeline.log 2>&1
tail -3 gpurun_out/r2_parity.log gpurun_out/r2_tn_batch_test.log gpurun_out/r2_interleaved.log gpurun_out/r2_pdl.log gpurun_out/r2_suite_ring8.log gpurun_out/r2_timeline.log
