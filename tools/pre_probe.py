"""Where the wall-clock of preprocess_subject goes (host staging, H2D, device work), cold and warm."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from multimodalsignal_b200 import preprocess as pp, synth

subs = [synth.make_subject(s, i, minutes=100.0) for i, s in enumerate(synth.ALL_SUBJECTS[:4])]
datas = [s.as_pickle_dict() for s in subs]
torch.cuda.init()
torch.zeros(1, device="cuda")
for rnd in range(3):
    t0 = time.perf_counter()
    for s, d in zip(subs, datas):
        t1 = time.perf_counter()
        out = pp.preprocess_subject(s.sid, d, pp.base_halving_quirk(s.sid, s.protocol), 64, include_wrist=True)
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"round {rnd} {s.sid}: enqueue {1e3*(t2-t1):7.1f} ms  + drain {1e3*(t3-t2):6.1f} ms")
    print(f"round {rnd}: {1e3*(time.perf_counter()-t0):.1f} ms for {len(subs)} subjects")
# the pieces, warm
chest = {k.decode(): v for k, v in datas[0][b"signal"][b"chest"].items()}
dev = torch.device("cuda", 0)
for _ in range(2):
    t0 = time.perf_counter(); rows = pp._UPLOADER.rows(chest, pp.CHEST_CHANNELS, dev); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"uploader.rows: host {1e3*(t1-t0):.1f} ms, then H2D drain {1e3*(t2-t1):.1f} ms ({rows.numel()*8/1e6:.0f} MB)")
num = pp.resampled_length(rows.shape[1], 700, 64)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); y = pp.resample_on_device(rows, num); torch.cuda.synchronize()
    print(f"resample_on_device 8 x {rows.shape[1]} -> {num}: {1e3*(time.perf_counter()-t0):.1f} ms")
# raw host copy speeds
big = np.random.rand(4_200_000, 3)
pin = torch.empty(big.size, dtype=torch.float64).pin_memory()
for _ in range(2):
    t0 = time.perf_counter(); np.copyto(pin.numpy().reshape(big.shape), big); print(f"contiguous memcpy 101 MB: {1e3*(time.perf_counter()-t0):.1f} ms")
    t0 = time.perf_counter(); np.copyto(pin.numpy()[:4_200_000], big[:, 0]); print(f"strided column copy 34 MB: {1e3*(time.perf_counter()-t0):.1f} ms")
d = torch.empty(big.size, dtype=torch.float64, device=dev)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pin, non_blocking=True); torch.cuda.synchronize(); print(f"H2D 101 MB pinned: {1e3*(time.perf_counter()-t0):.1f} ms")
t0 = time.perf_counter(); p2 = torch.empty(269_000_000 // 8, dtype=torch.float64).pin_memory(); print(f"pin_memory alloc 269 MB: {1e3*(time.perf_counter()-t0):.1f} ms")
