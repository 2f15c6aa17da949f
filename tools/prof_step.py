"""A few eager (no CUDA graph, side streams off) training steps at the headline configuration, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:<kernels> -s <n> -c <m> -o gpurun_out/x python tools/prof_step.py [steps]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodalsignal_b200 import _ext
from multimodalsignal_b200.models import CnnGruAttentionModel
from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, C, T = 64, 6, 3840
lib = _ext.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = CnnGruAttentionModel(C, 2, dropout=0.5).to(dev).train()
opt = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
step = FusedTrainStep(model, opt, B, T, use_graph=False)
lib.mms_set_side_streams(0)
x = torch.randn(4, B, C, T, device=dev)
y = torch.randint(0, 2, (4, B), device=dev)
for i in range(steps):
    step(x[i % 4], y[i % 4])
torch.cuda.synchronize()
print("ok", step.last_loss())
