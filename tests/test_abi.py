"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/mms_b200.h declares (no compute call is made here)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "mms_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mms_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from multimodalsignal_b200 import _ext
    from multimodalsignal_b200.build import build
    build()
    lib = _ext.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_ext.EXPORTED_SYMBOLS) == names
    assert lib.mms_version() >= 100


def test_param_layout_matches_state_dict():
    """Layout queries are host-only: every parameter has a 16-byte aligned slot, both GRU
    directions are adjacent, and the total equals the reference's parameter count + padding."""
    import warnings
    from multimodalsignal_b200.models import CnnGruAttentionModel
    for C, nc, kw, nparams in [(6, 2, {}, 123854), (3, 2, {}, 123506), (8, 2, {}, 124098), (14, 2, {}, 124822),
                               (8, 2, {"gru_hidden_size": 32, "gru_num_layers": 1}, None)]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = CnnGruAttentionModel(C, nc, **kw)
        layout, total = m.flat_layout()
        if nparams is not None:
            assert sum(p.numel() for p in m.parameters()) == nparams        # SURVEY §8a probe
        spans = sorted((off, off + n) for _, off, n, _ in layout if n)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0
        assert spans[-1][1] <= total and total % 4 == 0
        by = {name: off for name, off, n, _ in layout}
        n_ih = m.gru.weight_ih_l0.numel()
        assert by["gru.weight_ih_l0_reverse"] == by["gru.weight_ih_l0"] + n_ih
        assert len(m.state_dict()) == 34 if kw == {} else True


def test_no_cpu_fallback_without_gpu():
    import torch
    from multimodalsignal_b200 import _ext
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_ext.MmsError):
        _ext.lib()
