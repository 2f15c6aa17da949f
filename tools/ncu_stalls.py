"""Summarise an ``ncu --set full --import-source on`` report: per kernel launch, the SASS instructions with
the most warp-stall samples and the stall-reason totals.  Usage:
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-regex] [top-n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
    if kre:
        cmd += ["--kernel-name", f"regex:{kre}"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": line, "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    for bi, b in enumerate(blocks):
        rows = list(csv.reader(io.StringIO("\n".join(b["lines"]))))
        if not rows:
            continue
        hdr = rows[0]
        idx = {h: i for i, h in enumerate(hdr)}
        body = [r for r in rows[1:] if len(r) == len(hdr)]
        samp = idx["# Samples"] if "# Samples" in idx else idx["Warp Stall Sampling (All Samples)"]
        tot = sum(int(r[samp] or 0) for r in body)
        print(f"=== launch {bi}: {b['name'][:110]}  total samples {tot}, {len(body)} SASS instrs")
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        rs = {h: sum(int(r[idx[h]] or 0) for r in body) for h in reasons}
        print("   reasons:", ", ".join(f"{k[6:]}={v} ({100*v/max(1,tot):.0f}%)" for k, v in sorted(rs.items(), key=lambda kv: -kv[1]) if v))
        order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:top]
        for i in sorted(order):
            r = body[i]
            s = int(r[samp] or 0)
            why = sorted(((int(r[idx[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
            print(f"   {i:5d} {100*s/max(1,tot):5.1f}%  {r[idx['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")


if __name__ == "__main__":
    main()
