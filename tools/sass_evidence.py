"""Static evidence from the shipped library: per kernel, the SASS instruction count and the Blackwell-specific
mnemonics it contains (cuobjdump -sass; table in /opt/skills/guides/B200_PROFILING.md), plus ptxas' register /
spill summary from the build log.  Usage:  python tools/sass_evidence.py [out.txt]
`kernel_facts()` is what tests/test_build_static.py asserts on."""
from __future__ import annotations

import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "multimodalsignal_b200" / "libmms_b200.so"
LOG = ROOT / "multimodalsignal_b200" / "csrc" / "build.log"

PATTERNS = OrderedDict([
    ("UTMALDG (TMA load)", r"\bUTMALDG"),
    ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"),
    ("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"),
    ("UTCBAR/commit", r"\bUTCBAR"),
    ("LDTM (tcgen05.ld)", r"\bLDTM"),
    ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("FFMA2 (fma.rn.f32x2)", r"\bFFMA2"),
    ("*.SYS loads/stores (peer memory)", r"\.SYS\b"),
    ("RED (red.global.add)", r"\bRED\b|\bREDG"),
    ("MUFU.EX2/RCP", r"\bMUFU\.(EX2|RCP)\b"),
    ("LDG/STG .256 (256-bit global access)", r"\b(LDG|STG)\.[A-Z0-9.]*256"),
    ("DFMA/DMUL/DADD (float64)", r"\b(DFMA|DMUL|DADD)\b"),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"^void ", "", re.sub(r"\(.*$", "", o)).replace("mms::", "").replace("(int)", "") for o in out]


def kernel_facts():
    """{kernel name: {"instrs": n, mnemonic label: count, ...}} from the shared library."""
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    facts, cur, names = OrderedDict(), None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            names.append(cur)
            facts[cur] = {"instrs": 0}
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            continue
        facts[cur]["instrs"] += 1
        for label, pat in PATTERNS.items():
            if re.search(pat, line):
                facts[cur][label] = facts[cur].get(label, 0) + 1
    pretty = demangle(names)
    return OrderedDict(sorted(((p, facts[n]) for p, n in zip(pretty, names)), key=lambda kv: kv[0]))


def ptxas_facts():
    """{mangled entry: {"registers": r, "spill_stores": s, "spill_loads": l, "smem": b}} from the build log."""
    out, cur = {}, None
    for line in LOG.read_text().splitlines():
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = {}
            continue
        if cur:
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                out[cur].update(stack=int(m.group(1)), spill_stores=int(m.group(2)), spill_loads=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                out[cur]["registers"] = int(m.group(1))
                sm = re.search(r"(\d+) bytes smem", line)
                out[cur]["smem"] = int(sm.group(1)) if sm else 0
    return out


def main():
    facts = kernel_facts()
    lines = ["# SASS evidence (cuobjdump -sass multimodalsignal_b200/libmms_b200.so, sm_100a); tools/sass_evidence.py",
             "# per kernel: instruction count and the Blackwell-specific mnemonics it contains (B200_PROFILING.md table)", ""]
    for name, f in facts.items():
        extra = ", ".join(f"{k}: {v}" for k, v in f.items() if k != "instrs")
        lines.append(f"{name:<62} {f['instrs']:>5} instrs   {extra}")
    px = ptxas_facts()
    spills = {k: v for k, v in px.items() if v.get("spill_stores") or v.get("spill_loads")}
    lines += ["", f"# ptxas: {len(px)} entry functions, {len(spills)} with register spills"
              + ("" if not spills else ": " + ", ".join(spills))]
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
