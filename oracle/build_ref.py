"""Stage the UNMODIFIED reference modules under ``baseline/_ref/`` so they travel to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (17LiQi/MultimodalSignal) is pure Python: there is nothing
to compile and no ``setup.py`` / ``pyproject.toml`` for ``pip install --target`` to act on, so the "build" of
the reference arm is a byte-for-byte staging of the five modules on the hot path (models.py, dataset.py,
trainer.py, preprocess.py, main.py) from ``/root/reference`` into ``baseline/_ref/``.  That directory is
git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so ``bench.py --impl
reference`` and the ``library_gpu_baseline`` leg import the reference's own ``models.py`` on the GPU box
(``oracle/ref_harness.py`` looks there when ``/root/reference`` is absent).  A manifest with the sha256 of
every staged file is written beside them; ``ref_harness.staged_ok()`` re-checks it.

``__graft_entry__.build()`` calls ``stage()`` whenever ``/root/reference`` exists (authoring container).
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"
FILES = ("models.py", "dataset.py", "trainer.py", "preprocess.py", "main.py")


def stage(force: bool = False) -> Path | None:
    if not (SRC / "models.py").exists():
        return DST if (DST / "MANIFEST.json").exists() else None
    DST.mkdir(parents=True, exist_ok=True)
    manifest = {}
    for name in FILES:
        data = (SRC / name).read_bytes()
        manifest[name] = hashlib.sha256(data).hexdigest()
        out = DST / name
        if force or not out.exists() or out.read_bytes() != data:
            shutil.copyfile(SRC / name, out)
            out.chmod(0o644)
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "sha256": manifest}, indent=1))
    return DST


if __name__ == "__main__":
    print(stage(force=True))
