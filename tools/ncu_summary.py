"""Collect the per-launch facts bench.py and DESIGN.md quote from ``ncu --set full`` reports into one JSON:
    python tools/ncu_summary.py profiles/r1_ncu_full_summary.json gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep ...
For every kernel launch: duration, DRAM bytes read / written, grid, block, registers, tensor-pipe utilisation,
achieved occupancy, executed warp instructions and the share of stall samples by reason (top 4)."""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

WANT = {
    "duration_us": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
    "grid": "launch__grid_size", "block": "launch__block_size", "registers": "launch__registers_per_thread",
    "tensor_pipe_pct": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "warp_instructions": "smsp__inst_executed.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
}
SCALE = {"usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6, "s": 1e6,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    out = defaultdict(list)
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").replace("mms::", "")
            rec = {"file": rep.split("/")[-1].replace(".ncu-rep", "")}
            for key, metric in WANT.items():
                if metric in idx and r[idx[metric]] not in ("", "n/a"):
                    v = float(r[idx[metric]].replace(",", ""))
                    rec[key] = v * SCALE.get(units[idx[metric]], 1.0)
            out[short].append(rec)
    json.dump(out, open(out_path, "w"), indent=1)
    for k, v in out.items():
        print(k, len(v), [round(x.get("duration_us", 0), 1) for x in v])


if __name__ == "__main__":
    main()
