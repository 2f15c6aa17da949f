#!/bin/bash
mkdir -p gpurun_out
T=r2c33
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/${T}_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/${T}_suite.log
timeout 600 python tools/ab_variants.py --quick --steps 300 --out gpurun_out/${T}_ab_parity.json "GRU_BWD_LATE=1" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r=json.loads(ln)
    except Exception: continue
    print(r['options'], r.get('parity',{}).get('ok'), r.get('timing',{}).get('ms_per_step'), {k:v for k,v in r.get('timing',{}).get('avg_us',{}).items() if 'gru' in k or 'pool_fwd' in k})
"
timeout 600 python tools/ab_variants.py --interleave 4 --steps 300 --out gpurun_out/${T}_ab.json "GRU_BWD_LATE=0" "GRU_BWD_LATE=1" 2>&1 | tail -3
