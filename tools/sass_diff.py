"""Are the kernels of the working tree the ones a given commit shipped?  Builds that commit's csrc/ in a scratch directory
and compares every kernel's SASS (instruction text AND encodings, whitespace-normalised) with the in-tree objects.

    python tools/sass_diff.py <commit>          e.g. the last commit whose library ran on a GPU

Used at the end of round 1, after the GPU budget was spent, to show that the refactors made since (templated / included
TN GEMM body, launch macros, option plumbing) left every default-path kernel byte-identical: 77 identical, 0 different,
11 new (opt-in kernels and the timeline stamp) against 0157417.
"""
from __future__ import annotations

import glob
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr"]
RENAMED = {"tc_gemm_tn_kernelILi2EEEv14CUtensorMap_stS1_": "tc_gemm_tn_kernelE14CUtensorMap_stS0_"}     # templated on the stage count


def kernels(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            out[cur].append(" ".join(re.sub(r"^\s*/\*[0-9a-f]{4,5}\*/", "", line).split()))
        elif cur and re.match(r"\s+/\* 0x[0-9a-f]{16} \*/", line):
            out[cur].append(line.strip())
    return out


def main():
    commit = sys.argv[1]
    from multimodalsignal_b200.build import build
    build()
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", str(ROOT), "archive", commit, "multimodalsignal_b200/csrc", "include"], capture_output=True, check=True)
        subprocess.run(["tar", "x", "-C", tmp], input=tar.stdout, check=True)
        src = Path(tmp) / "multimodalsignal_b200" / "csrc"
        procs = [subprocess.Popen(["nvcc", *FLAGS, "-c", str(f), "-o", str(f.with_suffix(".o"))], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                 for f in sorted(src.glob("*.cu"))]
        for p in procs:
            p.wait()
        same, differ, new = 0, [], []
        for f in sorted(glob.glob(str(src / "*.o"))):
            mine = ROOT / "multimodalsignal_b200" / "csrc" / os.path.basename(f)
            if not mine.exists():
                continue
            a, b = kernels(f), kernels(str(mine))
            for k, code in b.items():
                ka = k if k in a else None
                for new_frag, old_frag in RENAMED.items():
                    if ka is None and new_frag in k and k.replace(new_frag, old_frag) in a:
                        ka = k.replace(new_frag, old_frag)
                if ka is None:
                    new.append(k)
                elif a[ka] == code:
                    same += 1
                else:
                    differ.append(k)
    print(f"against {commit}: {same} kernels identical (instruction text and encodings), {len(differ)} different, {len(new)} new")
    for k in differ:
        print("  different:", k)
    for k in new:
        print("  new:", k)
    return 1 if differ else 0


if __name__ == "__main__":
    sys.path.insert(0, str(ROOT))
    sys.exit(main())
