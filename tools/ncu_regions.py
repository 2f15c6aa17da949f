"""Warp-stall samples of one kernel of an ``ncu --set full --import-source on`` report, summed over buckets of SASS
instructions (where inside the kernel the time goes):  python tools/ncu_regions.py report.ncu-rep kernel-regex [bucket]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
rows = list(csv.reader(io.StringIO("\n".join(blocks[0]))))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) == len(hdr)]
samp = idx["# Samples"] if "# Samples" in idx else idx["Warp Stall Sampling (All Samples)"]
ex = idx.get("Instructions Executed")
tot = sum(int(r[samp] or 0) for r in body)
for b0 in range(0, len(body), bucket):
    rs = body[b0:b0 + bucket]
    s = sum(int(r[samp] or 0) for r in rs)
    e = sum(int(r[ex] or 0) for r in rs) if ex is not None else 0
    ops = {}
    for r in rs:
        op = r[idx["Source"]].split()[0] if r[idx["Source"]].split() else ""
        if op.startswith("@"):
            op = r[idx["Source"]].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    top = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print(f"{b0:5d}-{b0+len(rs)-1:5d}  samples {s:5d} ({100*s/max(1,tot):4.1f}%)  warp-instr executed {e:9d}   {top}")
