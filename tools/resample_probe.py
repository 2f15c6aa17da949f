"""One warm and one measured chest-sized resample call (8 signals, N = 4 200 959 -> 384 087); run under
``ncu --launch-skip 20 --launch-count 20 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`` for the
per-launch durations and DRAM traffic of the second call."""
import sys
import torch
sys.path.insert(0, ".")
from multimodalsignal_b200 import preprocess as pp

n = 4200000 + 137 * 7
num = pp.resampled_length(n, 700, 64)
x = torch.randn(8, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    y = pp.resample_on_device(x, num)
torch.cuda.synchronize()
print("done", y.shape)
