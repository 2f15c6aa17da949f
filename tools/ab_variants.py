"""A/B harness for the run-time kernel switches (mms_set_option): for every variant, parity of the
affected ops against the oracle / golden fixtures, then the time of the graph-replayed training step
(BASELINE configs[1]: B = 64, C = 6, T = 3840) and the live per-kernel averages of the recurrences.

    python tools/ab_variants.py [--out gpurun_out/ab.json] [--steps 300] NAME=V[,NAME=V...] ...

Each positional argument is one variant (a comma-separated list of option settings); the library's defaults
(no option set) are always measured first.  Results are appended to the output file after every
variant, so a run that is cut short still leaves what it measured.  One process, one GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def set_options(lib, settings):
    """None clears an option (back to the library's built-in default)."""
    for k, v in settings.items():
        rc = lib.mms_clear_option(k.encode()) if v is None else lib.mms_set_option(k.encode(), int(v))
        assert rc == 0, (k, v)


def parity(lib, quick=False):
    """GRU op cases (tests/test_gpu_ops.py) and whole-model gradients (tests/test_gpu_model.py)."""
    import test_gpu_ops as ops
    import test_gpu_model as tm
    t0 = time.perf_counter()
    done = []
    try:
        for case in [(5, 40, 64, 32, False), (5, 40, 64, 32, True), (3, 33, 32, 32, False), (3, 33, 32, 16, True),
                     (2, 240, 64, 128, False), (150, 12, 64, 32, False)]:
            ops._gru_case(lib, *case)
            done.append(f"gru{case}")
        ops._gru_case(lib, 4, 20, 64, 128, True, steps=1)
        ops._gru_case(lib, 4, 20, 64, 128, False, steps=3)
        ops._gru_case(lib, 4, 20, 64, 128, False, steps=5)
        done.append("gru_steps_1_3_5")
        for case in (["c6_t640"] if quick else ["c6_t640", "c8_h32_l1", "c3_t336_ternary"]):
            tm.test_train_forward_backward_vs_reference(case)
            tm.test_gradients_vs_float64_oracle(case)
            done.append(f"model_{case}")
        if not quick:
            tm.test_input_gradient_matches_oracle()
            done.append("input_gradient")
        import __graft_entry__ as entry      # B = 8, T = 3840: M = 1920 rows, the tcgen05 NT / TN GEMM paths, vs the oracle
        entry.smoke()
        done.append("smoke_tcgen05_sized")
        return {"ok": True, "checked": done, "seconds": round(time.perf_counter() - t0, 2)}
    except Exception as e:      # noqa: BLE001 - the harness reports, it does not judge
        return {"ok": False, "checked": done, "error": f"{type(e).__name__}: {str(e)[:400]}",
                "trace": traceback.format_exc()[-800:], "seconds": round(time.perf_counter() - t0, 2)}


def timing(lib, steps, warmup=30, B=64, Cc=6, T=3840):
    import torch
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = CnnGruAttentionModel(Cc, 2, dropout=0.5).to(dev).train()
    opt = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
    step = FusedTrainStep(model, opt, B, T, use_graph=True)
    NB = 48
    gen = torch.Generator(device=dev).manual_seed(1234)
    pool_x = torch.randn(NB, B, Cc, T, device=dev, generator=gen)
    pool_y = torch.randint(0, 2, (NB, B), device=dev, generator=gen)
    pool_x[:, :, 0, :] += pool_y[:, :, None].float() * 0.5
    for i in range(warmup):
        step(pool_x[i % NB], pool_y[i % NB])
    torch.cuda.synchronize()
    best = None
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(pool_x[i % NB], pool_y[i % NB])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        best = ms if best is None else min(best, ms)
    loss = float(step.last_loss())
    # live per-kernel averages (eager, serialised launches)
    PROF = 10
    lib.mms_set_side_streams(0)
    lib.mms_profile_enable(1)
    for i in range(PROF):
        step.load(pool_x[i % NB], pool_y[i % NB])
        step._enqueue()
    buf = (C.c_char * 16384)()
    lib.mms_profile_report(buf, 16384)
    lib.mms_profile_enable(0)
    lib.mms_set_side_streams(1)
    kern = {}
    for line in buf.value.decode().strip().splitlines():
        name, cnt, tot = line.rsplit(" ", 2)
        kern[name] = round(1e3 * float(tot) / int(cnt), 2)
    del step, opt, model, pool_x, pool_y
    torch.cuda.empty_cache()
    return {"ms_per_step": round(best, 5), "windows_per_s": round(B / best * 1e3, 1), "loss_after": loss,
            "avg_us": {k: kern[k] for k in sorted(kern)}}


def interleaved(lib, variants, steps, reps, B=64, Cc=6, T=3840):
    """Graph-replayed step of every variant, captured up front, timed in alternation (A B C A B C ...) so that clock and
    thermal drift hit all variants alike.  The pseudo-option SIDE_STREAMS (default 1) is mms_set_side_streams at capture
    time: 0 puts every kernel of the step on one stream, so the step time is the plain sum of the kernel times."""
    import torch
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep
    dev = torch.device("cuda", 0)
    NB = 48
    gen = torch.Generator(device=dev).manual_seed(1234)
    pool_x = torch.randn(NB, B, Cc, T, device=dev, generator=gen)
    pool_y = torch.randint(0, 2, (NB, B), device=dev, generator=gen)
    pool_x[:, :, 0, :] += pool_y[:, :, None].float() * 0.5
    built = []
    names = set().union(*[set(v) for v in variants]) - {"SIDE_STREAMS"}
    for settings in variants:
        opts = {k: (int(settings[k]) if k in settings else None) for k in names}      # options a variant does not name are cleared
        set_options(lib, opts)
        lib.mms_set_side_streams(int(settings.get("SIDE_STREAMS", 1)))
        torch.manual_seed(42)
        model = CnnGruAttentionModel(Cc, 2, dropout=0.5).to(dev).train()
        opt = FlatAdam(model, lr=1e-3, weight_decay=1e-4)
        step = FusedTrainStep(model, opt, B, T, use_graph=True)
        for i in range(20):
            step(pool_x[i % NB], pool_y[i % NB])
        torch.cuda.synchronize()
        built.append((settings, step, model, opt, []))
    lib.mms_set_side_streams(1)
    set_options(lib, {k: None for k in names})
    for rep in range(reps):
        for settings, step, _, _, times in built:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(pool_x[i % NB], pool_y[i % NB])
            e1.record()
            torch.cuda.synchronize()
            times.append(round(e0.elapsed_time(e1) / steps, 5))
    return [{"options": s, "ms_per_step_reps": t, "ms_per_step_min": min(t), "ms_per_step_median": sorted(t)[len(t) // 2]}
            for s, _, _, _, t in built]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "ab.json"))
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--quick", action="store_true", help="fewer model parity cases")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--interleave", type=int, default=0, metavar="REPS",
                    help="capture every variant first, then time them in alternation REPS times (no parity, no kernel profile)")
    ap.add_argument("variants", nargs="*")
    args = ap.parse_args()
    import torch
    from multimodalsignal_b200 import _ext
    lib = _ext.lib()
    out = Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    results = {"device": torch.cuda.get_device_name(0), "steps": args.steps, "variants": []}

    def flush():
        out.write_text(json.dumps(results, indent=1))

    if args.interleave:
        variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in args.variants]
        results["interleaved"] = interleaved(lib, variants, args.steps, args.interleave)
        flush()
        for r in results["interleaved"]:
            print(json.dumps(r), flush=True)
        return
    variants = [{}] + [dict(kv.split("=") for kv in v.split(",")) for v in args.variants]
    names = set().union(*[set(v) for v in variants])
    for settings in variants:
        full = {n: (int(settings[n]) if n in settings else None) for n in names}    # options a variant does not name are cleared
        set_options(lib, full)
        rec = {"options": full}
        t0 = time.perf_counter()
        if settings and not args.no_parity:
            rec["parity"] = parity(lib, quick=args.quick)
            flush()
        if not settings or args.no_parity or rec["parity"]["ok"]:
            try:
                rec["timing"] = timing(lib, args.steps)
            except Exception as e:      # noqa: BLE001
                rec["timing"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        rec["seconds"] = round(time.perf_counter() - t0, 2)
        results["variants"].append(rec)
        flush()
        print(json.dumps(rec), flush=True)
    set_options(lib, {n: None for n in names})


if __name__ == "__main__":
    main()
