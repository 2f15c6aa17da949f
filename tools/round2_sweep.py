"""First GPU call of round 2: everything the end of round 1 left unmeasured, ~4 minutes on one B200.

    gpurun --timeout 480 -- 'python tools/round2_sweep.py'

1. Every switch that selects a different KERNEL is checked for parity in its OWN process (a faulting kernel poisons the CUDA
   context): GRU op cases, golden model cases and smoke() at the tcgen05-sized batch against the oracle, then the step time
   and the eager per-kernel times (tools/ab_variants.py).  Op-level tests of the kernels written without GPU access.
2. The variants that passed, plus the host-only scheduling switches, as separately captured CUDA graphs timed in alternation
   against the default -- with the weight-gradient kernels on side streams and, for reference, on one stream.
3. The programmatic-dependent-launch build (libmms_b200_pdl.so): parity and step time.
4. The whole GPU suite with the depth-8 backward ring (green = it may become the default).
5. tools/graph_timeline.py: where every launch starts and ends INSIDE the replayed graph, side streams on and off.
Everything lands in gpurun_out/r2_*; a summary is printed at the end.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PY = sys.executable

KERNEL_VARIANTS = ["GRU_BWD_RING=8", "TN_STAGES=1", "TN_SPLIT=48", "WGRAD1_TILE=480", "WGRAD2_TILE=120", "NT_TRIM_STAGES=1",
                   "CONV_FWD_V2=1", "CONV_DGRAD_V2=1", "GRU_FWD_V2=1", "TN_BATCH=1", "TN_BATCH=1,TN_STAGES=1"]
SCHEDULING_VARIANTS = ["GRU_BWD_EXCLUSIVE_KB=200", "WGRAD_DEFER=1"]        # same kernels, different placement: no parity risk
COMBINATIONS = ["TN_STAGES=1,TN_SPLIT=48", "TN_STAGES=1,WGRAD1_TILE=480", "TN_STAGES=1,WGRAD1_TILE=480,WGRAD2_TILE=120", "TN_BATCH=1,WGRAD1_TILE=480",
                "GRU_FWD_V2=1,CONV_FWD_V2=1,CONV_DGRAD_V2=1,GRU_BWD_RING=8"]


def run(cmd, log, timeout, env=None):
    e = dict(os.environ)
    e.update(env or {})
    with open(OUT / log, "w") as f:
        try:
            return subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=timeout, env=e, cwd=ROOT).returncode
        except subprocess.TimeoutExpired:
            f.write(f"\nTIMEOUT after {timeout} s\n")
            return -1


def tag(v):
    return v.replace("=", "").replace(",", "_")


def main():
    OUT.mkdir(exist_ok=True)
    summary = {"parity": {}, "op_tests": {}}
    ok = []
    for v in KERNEL_VARIANTS:
        out = OUT / f"r2_parity_{tag(v)}.json"
        rc = run([PY, "tools/ab_variants.py", "--quick", "--out", str(out), v], f"r2_parity_{tag(v)}.log", 90)
        good = False
        if rc == 0 and out.exists():
            recs = json.loads(out.read_text())["variants"]
            good = len(recs) == 2 and recs[1].get("parity", {}).get("ok", False) and "ms_per_step" in recs[1].get("timing", {})
            if good:
                summary["parity"][v] = {"ok": True, "ms_per_step": recs[1]["timing"]["ms_per_step"],
                                        "default_ms_per_step": recs[0]["timing"]["ms_per_step"]}
        if good:
            ok.append(v)
        else:
            summary["parity"][v] = {"ok": False, "rc": rc, "log": f"gpurun_out/r2_parity_{tag(v)}.log"}
    summary["op_tests"]["tn_batch"] = run([PY, "-m", "pytest", "tests/test_gpu_tc_gemm.py", "-q", "-k", "batch"], "r2_optest_tn_batch.log", 90,
                                          {"MMS_TEST_EXPERIMENTAL": "1"})
    summary["op_tests"]["conv_v2"] = run([PY, "-m", "pytest", "tests/test_gpu_ops.py", "-q", "-k", "conv1d"], "r2_optest_conv_v2.log", 90,
                                         {"MMS_CONV_FWD_V2": "1", "MMS_CONV_DGRAD_V2": "1"})
    summary["op_tests"]["gru_fwd_v2"] = run([PY, "-m", "pytest", "tests/test_gpu_ops.py", "-q", "-k", "gru"], "r2_optest_gru_fwd_v2.log", 90,
                                            {"MMS_GRU_FWD_V2": "1"})

    def usable(combo):          # every part passed parity on its own
        return all(part in ok for part in combo.split(","))
    variants = ["GRU_BWD_RING=4"] + ok + SCHEDULING_VARIANTS + [c for c in COMBINATIONS if usable(c)] + ["SIDE_STREAMS=0,GRU_BWD_RING=4"]
    rc = run([PY, "tools/ab_variants.py", "--interleave", "3", "--steps", "300", "--out", str(OUT / "r2_interleaved.json"), *variants],
             "r2_interleaved.log", 240)
    summary["interleaved_rc"] = rc
    if (OUT / "r2_interleaved.json").exists():
        inter = json.loads((OUT / "r2_interleaved.json").read_text()).get("interleaved", [])
        summary["interleaved_ms_min"] = {",".join(f"{k}={v}" for k, v in r["options"].items()): r["ms_per_step_min"] for r in inter}
    # capture priority is a Python-side switch (read when the capture stream is created): its own process
    run([PY, "tools/ab_variants.py", "--interleave", "3", "--steps", "300", "--out", str(OUT / "r2_priority.json"), "GRU_BWD_RING=4"],
        "r2_priority.log", 90, {"MMS_CAPTURE_PRIORITY": "-1"})
    pdl = ROOT / "multimodalsignal_b200" / "libmms_b200_pdl.so"
    summary["pdl_rc"] = run([PY, "tools/ab_variants.py", "--quick", "--out", str(OUT / "r2_pdl.json"), "GRU_BWD_RING=4"], "r2_pdl.log", 90,
                            {"MMS_B200_LIB": str(pdl)})
    summary["suite_ring8_rc"] = run([PY, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q"], "r2_suite_ring8.log", 120, {"MMS_GRU_BWD_RING": "8"})
    summary["timeline_rc"] = run([PY, "tools/graph_timeline.py", "--out", str(OUT / "r2_timeline.json")], "r2_timeline.log", 90)
    (OUT / "r2_summary.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
