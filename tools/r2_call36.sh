#!/bin/bash
mkdir -p gpurun_out
T=r2c36
timeout 600 python tools/ab_variants.py --quick --steps 200 --out gpurun_out/${T}_ab_parity.json "NT_BN_CAP=96" "NT_BN_CAP=64" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r=json.loads(ln)
    except Exception: continue
    print(r['options'], r.get('parity',{}).get('ok'), r.get('parity',{}).get('error'), r.get('timing',{}).get('ms_per_step'), {k:v for k,v in r.get('timing',{}).get('avg_us',{}).items() if 'nt_k' in k})
"
timeout 600 python tools/ab_variants.py --interleave 4 --steps 300 --out gpurun_out/${T}_ab.json "NT_BN_CAP=256" "NT_BN_CAP=128" "NT_BN_CAP=96" "NT_BN_CAP=64" 2>&1 | tail -4
