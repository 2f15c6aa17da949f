"""Import the UNMODIFIED reference modules from ``/root/reference`` (authoring container) or,
on the GPU box where that tree does not exist, from the byte-identical staging ``baseline/_ref``
made by ``oracle/build_ref.py`` (git-ignored; it travels with the gpurun snapshot).

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` to produce the
committed fixtures under ``tests/golden/`` and by ``bench.py --impl reference``
when the reference tree is present.

The reference imports three packages that are absent here and are not on the hot
path: ``matplotlib`` / ``seaborn`` (trainer.py:9-10, plotting only -- the plotting
call is already wrapped in try/except, trainer.py:250-273) and ``neurokit2``
(preprocess.py:9, feature branch only).  Empty stub modules are injected for
those names; nothing else is changed.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types
from pathlib import Path

_STAGED = Path(__file__).resolve().parents[1] / "baseline" / "_ref"     # oracle/build_ref.py (git-ignored, travels to the GPU box)


def _find_root() -> Path:
    env = os.environ.get("MMS_REFERENCE_ROOT")
    if env:
        return Path(env)
    if Path("/root/reference/models.py").exists():
        return Path("/root/reference")
    return _STAGED


REFERENCE_ROOT = _find_root()
_STUBS = ("matplotlib", "matplotlib.pyplot", "seaborn", "neurokit2")


def available() -> bool:
    return (REFERENCE_ROOT / "models.py").exists()


def staged_ok() -> bool:
    """True when the files under ``baseline/_ref`` still hash to their manifest (i.e. are the unmodified reference)."""
    import hashlib
    import json
    man = _STAGED / "MANIFEST.json"
    if not man.exists():
        return False
    want = json.loads(man.read_text())["sha256"]
    return all((_STAGED / n).exists() and hashlib.sha256((_STAGED / n).read_bytes()).hexdigest() == h for n, h in want.items())


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load(name: str, scratch_dir=None):
    """Import reference module ``name`` under the alias ``mms_ref_<name>`` so it can
    never shadow (or be shadowed by) this repo's modules of the same name.
    ``preprocess`` creates ``./data`` directories at import time (preprocess.py:14-15),
    so it is imported with ``scratch_dir`` as the working directory."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    alias = f"mms_ref_{name}"
    if alias in sys.modules:
        return sys.modules[alias]
    # the reference modules import each other by bare name (main.py:9-11)
    saved = {k: sys.modules.get(k) for k in ("models", "dataset", "trainer", "preprocess", "main")}
    sys.path.insert(0, str(REFERENCE_ROOT))
    try:
        for k in saved:
            sys.modules.pop(k, None)
        with _cwd(scratch_dir or os.getcwd()):
            mod = importlib.import_module(name)
        sys.modules[alias] = mod
        return mod
    finally:
        sys.path.remove(str(REFERENCE_ROOT))
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
