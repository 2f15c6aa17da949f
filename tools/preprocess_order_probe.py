"""15 subjects of different lengths through resample_subject_rows + window gather, three passes, per pass wall/device time,
for the three enqueue orders of preprocess.resample_many (MMS_RESAMPLE_MANY_ORDER)."""
import sys, time
import torch
sys.path.insert(0, ".")
from multimodalsignal_b200 import preprocess as pp

torch.manual_seed(0)
subs = []
for i in range(15):
    n = 4200000 + 9973 * i + 137
    rows = torch.randn(8, n, dtype=torch.float64, device="cuda")
    wr = {name: torch.randn(3 if name == "ACC" else 1, int(n * fs / 700), dtype=torch.float64, device="cuda") for name, fs in pp.WRIST_CHANNELS.items()}
    subs.append((rows, wr))
torch.cuda.synchronize()
for order in (1, 0, 1):
    pp._MANY_ORDER = order
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for rows, wr in subs:
            s = pp.resample_subject_rows(rows, wr, 64)
            del s
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        print(f"order {order} rep {rep}: {1e3 * t / 15:.3f} ms per subject (host enqueue {1e3 * t_host / 15:.3f})", flush=True)
