#!/bin/bash
mkdir -p gpurun_out
for raw in 0 1; do
echo "== TF32_RAW_HI=$raw"
MMS_TF32_RAW_HI=$raw timeout 300 python - <<'PY'
import torch, ctypes as C
from multimodalsignal_b200 import _ext
lib = _ext.lib()
torch.manual_seed(0)
for (M, N, K) in [(2048, 192, 128), (1920, 32, 512), (4096, 128, 192)]:
    A = torch.randn(M, K, dtype=torch.float64) * (1 + torch.rand(M, K, dtype=torch.float64))
    W = torch.randn(N, K, dtype=torch.float64)
    Ad, Wd = A.float().cuda(), W.float().cuda()
    Cd = torch.empty(M, N, device="cuda")
    _ext.check(lib.mms_tc_gemm_nt(Ad.data_ptr(), K, Wd.data_ptr(), K, None, Cd.data_ptr(), N, M, N, K, 0, torch.cuda.current_stream().cuda_stream))
    ref = Ad.double().cpu() @ Wd.double().cpu().t()
    err = (Cd.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    print(f"  {M}x{N}x{K}: max rel err {err:.3e}")
PY
done
timeout 600 python tools/ab_variants.py --interleave 4 --steps 300 --out gpurun_out/r2c44_ab.json "TF32_RAW_HI=0" "TF32_RAW_HI=1" 2>&1 | tail -2
