#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2i_bench8.json 2> gpurun_out/r2i_bench8.err ) 2> gpurun_out/r2i_time8.txt; echo "bench8 rc=$?"; grep real gpurun_out/r2i_time8.txt
tail -3 gpurun_out/r2i_bench8.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2i_bench8.json'))
print('value', round(p['value']), 'ms', round(p['ms_per_step'],4), 'e2e', round(p['e2e']['value']))
for k in ('loso','dp'):
    v=p.get(k,{})
    print(k, {a:b for a,b in v.items() if a in ('value','unit','unavailable','preprocess_s','train_windows_per_s','accuracy_mean','ms_per_step','nccl','windows_trained')})
PY
