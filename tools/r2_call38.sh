#!/bin/bash
timeout 300 python - <<'PY' 2>&1 | grep "^NT" | tail -8
import torch, sys
sys.path.insert(0, ".")
from multimodalsignal_b200.models import CnnGruAttentionModel
from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep
dev = torch.device("cuda", 0)
torch.manual_seed(42)
B, Cc, T = 64, 6, 3840
model = CnnGruAttentionModel(Cc, 2, dropout=0.5).to(dev).train()
opt = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
step = FusedTrainStep(model, opt, B, T, use_graph=False)
x = torch.randn(B, Cc, T, device=dev); y = torch.randint(0, 2, (B,), device=dev)
for i in range(3):
    step(x, y)
torch.cuda.synchronize()
PY
