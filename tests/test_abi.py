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


def test_option_registry_is_host_only_and_defaults_belong_to_the_caller(monkeypatch):
    """mms_set_option / mms_get_option / mms_clear_option (include/mms_b200.h): explicit value > environment (read once) >
    the default of the call site, which is never cached."""
    from multimodalsignal_b200 import _ext
    lib = _ext.load_library()
    assert lib.mms_get_option(b"TEST_ONLY_A", 7) == 7
    assert lib.mms_get_option(b"TEST_ONLY_A", 9) == 9          # a default passed earlier is not remembered
    assert lib.mms_set_option(b"TEST_ONLY_A", 3) == 0
    assert lib.mms_get_option(b"TEST_ONLY_A", 9) == 3
    assert lib.mms_clear_option(b"TEST_ONLY_A") == 0
    assert lib.mms_get_option(b"TEST_ONLY_A", 9) == 9
    monkeypatch.setenv("MMS_TEST_ONLY_B", "5")
    assert lib.mms_get_option(b"TEST_ONLY_B", 1) == 5          # environment, read at first use
    monkeypatch.setenv("MMS_TEST_ONLY_B", "6")
    assert lib.mms_get_option(b"TEST_ONLY_B", 1) == 5          # ... and only then
    assert lib.mms_set_option(b"TEST_ONLY_B", 2) == 0
    assert lib.mms_get_option(b"TEST_ONLY_B", 1) == 2
    assert lib.mms_set_option(b"", 1) < 0 and b"empty name" in lib.mms_last_error()
