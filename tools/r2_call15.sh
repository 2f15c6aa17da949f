#!/bin/bash
# N = 2: the driver's multi-GPU launch of the default bench line (loso + dp sub-records), then the 2-GPU tests
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/r2c15_bench2.json 2> gpurun_out/r2c15_bench2.err ) 2> gpurun_out/r2c15_time.txt; echo "bench2 rc=$?"; tail -3 gpurun_out/r2c15_time.txt
tail -5 gpurun_out/r2c15_bench2.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c15_bench2.json'))
print('value', p['value'], 'ms', p['ms_per_step'], 'e2e', p['e2e']['value'])
for k in ('loso','preprocess','dp'):
    v=p.get(k,{})
    print(k, {a:b for a,b in v.items() if a in ('value','unit','unavailable','preprocess_s','train_windows_per_s','accuracy_mean','ms_per_step','nccl')})
PY
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q -x > gpurun_out/r2c15_par.log 2>&1; echo "parallel tests rc=$?"; tail -4 gpurun_out/r2c15_par.log
