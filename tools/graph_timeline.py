"""In-graph timeline of the training step (mms_timeline_enable): where every kernel of the main chain and of the
weight-gradient side streams starts and ends INSIDE the replayed CUDA graph, with the side streams on and off.

    python tools/graph_timeline.py [--out gpurun_out/timeline.json] [NAME=V ...]      (options as in ab_variants.py)

The stamps (one-thread %globaltimer kernels around every launch) add about two launch latencies per kernel, so the
absolute step time is longer than the benchmark's; what the table is for is the overlap structure and the STRETCH of
a kernel when it runs beside others: duration with side streams / duration on a single stream.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def timeline(lib, side_streams, B=64, Cc=6, T=3840, replays=20):
    import torch
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep, capture_graph
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = CnnGruAttentionModel(Cc, 2, dropout=0.5).to(dev).train()
    opt = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
    step = FusedTrainStep(model, opt, B, T, use_graph=True)
    gen = torch.Generator(device=dev).manual_seed(1234)
    x = torch.randn(8, B, Cc, T, device=dev, generator=gen)
    y = torch.randint(0, 2, (8, B), device=dev, generator=gen)
    lib.mms_set_side_streams(int(side_streams))
    step(x[0], y[0])                                   # eager: kernel attributes, caches (timeline still off)
    torch.cuda.synchronize()
    assert lib.mms_timeline_enable(1) == 0
    step.load(x[1], y[1])
    graph = capture_graph(step._enqueue)               # the stamps are captured with the step
    for i in range(replays):
        step.load(x[i % 8], y[i % 8])
        graph.replay()
    buf = (C.c_char * (1 << 17))()
    assert lib.mms_timeline_report(buf, 1 << 17) == 0
    lib.mms_timeline_enable(0)
    lib.mms_set_side_streams(1)
    rows = []
    for line in buf.value.decode().strip().splitlines():
        name, sid, t0, t1 = line.rsplit(" ", 3)
        rows.append({"kernel": name, "stream": int(sid), "start_us": int(t0) / 1e3, "end_us": int(t1) / 1e3})
    rows.sort(key=lambda r: r["start_us"])
    del graph, step
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "timeline.json"))
    ap.add_argument("options", nargs="*")
    args = ap.parse_args()
    from multimodalsignal_b200 import _ext
    lib = _ext.lib()
    for kv in args.options:
        k, v = kv.split("=")
        assert lib.mms_set_option(k.encode(), int(v)) == 0
    on, off = timeline(lib, 1), timeline(lib, 0)
    # the k-th launch of a kernel name (by start time) is the same launch in both captures: launches of one name are ordered by
    # data dependencies in the step (e.g. gru_bwd#0 = top layer, gru_bwd#1 = layer 0)
    def by_order(rows_by_start):
        seen, out = {}, []
        for r in rows_by_start:
            i = seen.get(r["kernel"], 0)
            seen[r["kernel"]] = i + 1
            out.append((f'{r["kernel"]}#{i}', r))
        return out
    single = {k: r["end_us"] - r["start_us"] for k, r in by_order(off)}
    table = []
    for k, r in by_order(on):
        d = r["end_us"] - r["start_us"]
        table.append({"launch": k, "stream": r["stream"], "start_us": round(r["start_us"], 1), "end_us": round(r["end_us"], 1),
                      "us": round(d, 1), "us_single_stream": round(single.get(k, float("nan")), 1),
                      "stretch": round(d / single[k], 2) if single.get(k) else None})
    span_on = max(r["end_us"] for r in on) - min(r["start_us"] for r in on)
    span_off = max(r["end_us"] for r in off) - min(r["start_us"] for r in off)
    out = {"options": args.options, "span_us_side_streams": round(span_on, 1), "span_us_single_stream": round(span_off, 1),
           "note": "stamped graphs: spans include ~2 extra launches per kernel; compare structure and stretch, not absolutes",
           "launches": table}
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))
    print(f"span with side streams {span_on:.0f} us, single stream {span_off:.0f} us")
    print(f'{"launch":<28}{"stream":>7}{"start":>9}{"end":>9}{"us":>8}{"single":>8}{"stretch":>8}')
    for t in table:
        print(f'{t["launch"]:<28}{t["stream"]:>7}{t["start_us"]:>9}{t["end_us"]:>9}{t["us"]:>8}{t["us_single_stream"]:>8}{str(t["stretch"]):>8}')


if __name__ == "__main__":
    main()
