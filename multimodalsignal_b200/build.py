"""Build libmms_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

``python -m multimodalsignal_b200.build`` or ``build()`` from ``__graft_entry__``.
The shared library is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmms_b200.so"
STAMP = PKG / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v", *os.environ.get("MMS_NVCC_EXTRA", "").split(),
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) + [PKG.parent / "include" / "mms_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in _sources():
        obj = src.with_suffix(".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            failed = True
        objs.append(str(obj))
    (CSRC / "build.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see multimodalsignal_b200/csrc/build.log")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
    subprocess.run(link, check=True)
    STAMP.write_text(digest)
    if verbose:
        print("\n".join(log))
    return LIB


def build_pdl(force: bool = False) -> Path:
    """The programmatic-dependent-launch variant (-DMMS_PDL, see csrc/mms_common.cuh) as a SECOND library,
    ``libmms_b200_pdl.so``, beside the default one; select it with ``MMS_B200_LIB=<path>``.  Round-2 experiment:
    compiled and checked for the griddepcontrol instructions (ACQBULK / PREEXIT in the SASS), parity-green on a B200 in round 2 but slower than the default library there."""
    lib = PKG / "libmms_b200_pdl.so"
    stamp = CSRC / ".build_stamp_pdl"
    digest = _digest() + ":pdl"
    if not force and lib.exists() and stamp.exists() and stamp.read_text() == digest:
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = CSRC / "pdl_objs"
    objdir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")]
        cmd = [nvcc, *flags, "-DMMS_PDL", "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc -DMMS_PDL failed on {src.name}")
        objs.append(str(obj))
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *objs], check=True)
    stamp.write_text(digest)
    return lib


if __name__ == "__main__":
    if "--pdl" in sys.argv:
        print(build_pdl(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
