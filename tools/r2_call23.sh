#!/bin/bash
mkdir -p gpurun_out
T=r2c23
timeout 600 python tools/ab_variants.py --interleave 3 --steps 300 --out gpurun_out/${T}_ab.json "WGRAD_DEFER=1" "WGRAD_DEFER=2" "WGRAD_DEFER=2,GRU_BWD_EXCLUSIVE_KB=200" "WGRAD_DEFER=2,TN_STAGES=1" "WGRAD_DEFER=1,GEMM_SKINNY=0" "WGRAD_DEFER=2,GRU_BWD_EXCLUSIVE_KB=200,TN_STAGES=3" 2>&1 | tail -8
