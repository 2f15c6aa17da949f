#!/bin/bash
# 2-GPU pass: the data-parallel / sharded tests that need two devices, the GEMM tests with skip reasons, and the default bench line at N = 2
mkdir -p gpurun_out
T=r2i
timeout 900 python -m pytest tests/test_gpu_parallel.py tests/test_gpu_preprocess.py tests/test_gpu_tc_gemm.py -m gpu -q -rs > gpurun_out/${T}_parallel_tests_2gpu.log 2>&1; echo "2gpu tests rc=$?"; tail -12 gpurun_out/${T}_parallel_tests_2gpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err ) 2> gpurun_out/${T}_time2.txt; echo "bench2 rc=$?"; tail -3 gpurun_out/${T}_time2.txt
python - <<'PY'
import json
p=json.loads([l for l in open('gpurun_out/r2i_bench_2gpu.json') if l.startswith('{')][-1])
print({k:p[k] for k in ('value','ms_per_step','n_gpus') if k in p}, p['e2e'])
for k in ('loso','dp','preprocess'):
    if k in p:
        q=p[k]; print(' ',k,{kk:q[kk] for kk in ('value','unit','ms_per_step','unavailable','nccl') if kk in q})
PY
