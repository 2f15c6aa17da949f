"""Pin ``oracle/preprocess_oracle.py`` to the reference: ``scipy.signal.resample``
(the third-party routine the reference calls, preprocess.py:72,74) and fixtures made
by running the unmodified ``preprocess.py`` / ``dataset.py`` (oracle/make_golden.py)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from multimodalsignal_b200 import synth
from oracle import preprocess_oracle as po


def test_resample_matches_reference_cases():
    z, _ = load_golden("resample_cases.npz")
    for i in range(int(z["n_cases"])):
        x, y = z[f"x{i}"], z[f"y{i}"]
        f0, f1 = [int(v) for v in z[f"fs{i}"]]
        got = po.resample_signal(x, f0, f1)
        assert got.shape == y.shape and got.dtype == np.float64
        assert np.abs(got - y).max() <= 1e-12 * max(1.0, np.abs(y).max()), i


@pytest.mark.parametrize("n,num", [(700, 64), (701, 64), (64, 700), (65, 701), (128, 128), (1000, 500), (999, 333)])
def test_resample_matches_scipy(n, num):
    import scipy.signal
    x = np.random.default_rng(n).standard_normal(n)
    np.testing.assert_allclose(po.fft_resample(x, num), scipy.signal.resample(x, num), atol=1e-13)


def test_float_trap_values_are_kept():
    """int(32.05*60*700) is 1346099 in float64, not 1346100."""
    assert int(32.05 * 60 * 700) == 1346099
    starts, labels, w = po.window_plan([("TSST", 32.05, 43.5)], 64)
    assert w == 3840
    assert starts[0] == int(1346099 * (64 / 700))


@pytest.mark.parametrize("fs", [64, 128])
def test_window_plan_matches_reference_full_protocol(fs):
    z, _ = load_golden("preprocess_golden.npz")
    for sid in synth.ALL_SUBJECTS:
        starts, labels, w = po.window_plan(po.apply_subject_quirk(sid, synth.FULL_PROTOCOL), fs)
        assert np.array_equal(starts, z[f"full/{fs}/{sid}/starts"]), sid
        assert np.array_equal(labels, z[f"full/{fs}/{sid}/labels"]), sid
        assert tuple(z[f"full/{fs}/{sid}/shape"]) == (len(starts), w, 8)
    assert bytes(z["channel_names"]).decode().split() == [
        "chest_ACC_x", "chest_ACC_y", "chest_ACC_z", "chest_ECG", "chest_EDA", "chest_EMG", "chest_Resp", "chest_Temp"]


@pytest.mark.parametrize("fs", [64, 128])
@pytest.mark.parametrize("sid", ["S2", "S5"])
def test_preprocess_subject_matches_reference_short(fs, sid):
    z, meta = load_golden("preprocess_golden.npz")
    idx = synth.ALL_SUBJECTS.index(sid)
    sub = synth.make_subject(sid, idx, minutes=synth.SHORT_MINUTES, protocol=synth.SHORT_PROTOCOL, with_wrist=False)
    X, y = po.preprocess_subject(sid, sub.chest, sub.protocol, fs)
    assert tuple(z[f"short/{fs}/{sid}/shape"]) == X.shape
    assert np.array_equal(y, z[f"short/{fs}/{sid}/labels"])
    np.testing.assert_allclose(X[:, ::61, :], z[f"short/{fs}/{sid}/X_sub"], atol=1e-11)
    np.testing.assert_allclose(X.sum(axis=1), z[f"short/{fs}/{sid}/X_rowsum"], atol=1e-8)


@pytest.mark.parametrize("mode", ["stress_binary", "ternary"])
def test_dataset_normalisation_matches_reference(mode):
    z, _ = load_golden("preprocess_golden.npz")
    names = bytes(z["channel_names"]).decode().split()
    chans = ["chest_ECG", "chest_EDA", "chest_EMG", "chest_Resp"]
    data, labels = [], []
    for sid in ("S2", "S5"):
        idx = synth.ALL_SUBJECTS.index(sid)
        sub = synth.make_subject(sid, idx, minutes=synth.SHORT_MINUTES, protocol=synth.SHORT_PROTOCOL, with_wrist=False)
        X, y = po.preprocess_subject(sid, sub.chest, sub.protocol, 64)
        sel = X[:, :, [names.index(c) for c in chans]]
        data.append(po.normalise_subject(sel, chans))
        labels.append(po.map_labels(y, mode))
    data, labels = np.concatenate(data), np.concatenate(labels)
    assert len(labels) == int(z[f"dataset/{mode}/len"])
    assert np.array_equal(labels, z[f"dataset/{mode}/labels"])
    np.testing.assert_allclose(data[:, ::61, :], z[f"dataset/{mode}/data_sub"], atol=1e-9)
    item = data[3].astype(np.float32).T          # dataset.py:63 -> [C, W] float32
    np.testing.assert_allclose(item[:, ::61], z[f"dataset/{mode}/item3_x_sub"], atol=1e-6)
    assert int(labels[3]) == int(z[f"dataset/{mode}/item3_y"])
    with pytest.raises(ValueError):
        po.map_labels(np.array([1, 2]), "amusement_binary")


def test_fold_table_survey_values():
    """SURVEY §8d LOSO fold table (sklearn train_test_split, random_state=42)."""
    table = json.loads((GOLDEN / "fold_table.json").read_text())["folds"]
    assert table["S2"]["val"] == ["S13", "S15", "S3"]
    assert table["S2"]["train"] == ["S16", "S8", "S11", "S5", "S4", "S17", "S7", "S10", "S14", "S6", "S9"]
    for s in ["S3", "S4", "S5", "S6", "S7", "S8", "S9", "S10", "S11"]:
        assert table[s]["val"] == ["S13", "S15", "S2"]
    for s in ["S13", "S14"]:
        assert table[s]["val"] == ["S11", "S15", "S2"]
    for s in ["S15", "S16", "S17"]:
        assert table[s]["val"] == ["S11", "S14", "S2"]
