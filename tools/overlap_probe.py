"""How well do S co-resident training steps overlap on one B200?  Runs S independent (model, FusedTrainStep)
pairs, each on its own CUDA stream, reports whole-GPU windows/s for S = 1..Smax, and (with --trace) writes a
Kineto trace of a few steps so that per-kernel concurrency can be read offline.
    PYTHONPATH=. python tools/overlap_probe.py --smax 4 --trace gpurun_out/overlap_trace.json
"""
import argparse
import json
import time

import torch

from multimodalsignal_b200.models import CnnGruAttentionModel
from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--smax", type=int, default=4)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--trace", default="")
    ap.add_argument("--trace-streams", type=int, default=2)
    args = ap.parse_args()
    B, Cc, T = 64, 6, 3840
    dev = torch.device("cuda")
    pairs = []
    for s in range(args.smax):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            torch.manual_seed(s)
            m = CnnGruAttentionModel(Cc, 2).to(dev).train()
            step = FusedTrainStep(m, FlatAdam(m, lr=1e-3, weight_decay=1e-4), B, T)
            step.x.normal_()
            step.y.random_(0, 2)
            for _ in range(3):
                step.run()
        pairs.append((st, step))
    torch.cuda.synchronize()
    out = {}
    for S in range(1, args.smax + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for st, step in pairs[:S]:
                with torch.cuda.stream(st):
                    step.run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[S] = {"windows_per_s": S * args.steps * B / dt, "ms_per_round": 1e3 * dt / args.steps}
        print(f"S={S}: {out[S]['windows_per_s']:.0f} windows/s, {out[S]['ms_per_round']:.3f} ms per round of {S} steps", flush=True)
    if args.trace:
        from torch.profiler import ProfilerActivity, profile
        S = args.trace_streams
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(6):
                for st, step in pairs[:S]:
                    with torch.cuda.stream(st):
                        step.run()
            torch.cuda.synchronize()
        prof.export_chrome_trace(args.trace)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
