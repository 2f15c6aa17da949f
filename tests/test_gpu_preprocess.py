"""Preprocess path on a B200: FFT resampling (scipy.signal.resample semantics), window stacking and
the device-resident normalised dataset, against fixtures produced by the unmodified reference
(tests/golden/resample_cases.npz, preprocess_golden.npz) and against the numpy oracle.

Stated tolerance for the float64 resampler: |err| <= 1e-10 * max(1, max|y|) (the chirp-z evaluation
and pocketfft's mixed-radix/Bluestein evaluation of the same exact transform differ by float64
rounding only).  Window start indices, labels and gathered samples are bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from multimodalsignal_b200 import synth
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu

RESAMPLE_RTOL = 1e-10


def err_ok(got, ref, tol=RESAMPLE_RTOL):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(1.0, np.abs(ref).max())
    err = np.abs(got - ref).max()
    assert err <= tol * scale, f"max err {err:.3e} (scale {scale:.3e})"


def test_resample_signal_matches_reference_cases():
    from multimodalsignal_b200 import preprocess as pp
    z, _ = load_golden("resample_cases.npz")
    for i in range(int(z["n_cases"])):
        x, y = z[f"x{i}"], z[f"y{i}"]
        f0, f1 = [int(v) for v in z[f"fs{i}"]]
        got = pp.resample_signal(x, f0, f1)
        assert got.dtype == np.float64
        err_ok(got, y)


@pytest.mark.parametrize("n,num", [(2, 1), (3, 7), (17, 17), (700, 64), (701, 64), (64, 700), (1000, 500), (999, 333),
                                   (4096, 4096), (65537, 5991), (100000, 9142), (24000, 384000), (192015, 384030)])
def test_resample_matches_scipy_semantics(n, num):
    from multimodalsignal_b200 import preprocess as pp
    rng = np.random.default_rng(n + num)
    x = torch.from_numpy(rng.standard_normal((3, n)) + 0.5).cuda()
    got = pp.resample_on_device(x, num).cpu().numpy()
    ref = np.stack([po.fft_resample(r, num) for r in x.cpu().numpy()])
    err_ok(got, ref)


@pytest.mark.parametrize("n_sig", [2, 5, 8])
def test_resample_paired_transforms_equal_single_transforms(n_sig):
    """csrc/resample.cu: two real signals per complex chirp-z transform (the default when it fits) against one transform per
    signal (MMS_RESAMPLE_PAIRED=0) and against the numpy restatement of scipy.signal.resample; odd counts leave the last
    signal alone in its pair."""
    from multimodalsignal_b200 import _ext, preprocess as pp
    lib = _ext.lib()
    n, num = 70000 + 37, 6403
    rng = np.random.default_rng(n_sig)
    x = rng.standard_normal((n_sig, n)) * np.arange(1, n_sig + 1)[:, None] + np.arange(n_sig)[:, None]
    xd = torch.from_numpy(x).cuda()
    ref = np.stack([po.fft_resample(r, num) for r in x])
    prev = lib.mms_get_option(b"RESAMPLE_PAIRED", -1)
    try:
        outs = {}
        for mode in (1, 0):
            _ext.check(lib.mms_set_option(b"RESAMPLE_PAIRED", mode))
            outs[mode] = pp.resample_on_device(xd, num).cpu().numpy()
            err_ok(outs[mode], ref)
        assert np.abs(outs[1] - outs[0]).max() <= 1e-10 * max(1.0, np.abs(ref).max())
        # paired stage A with one inverse transform per signal (MMS_RESAMPLE_PAIRED_C=0) against the paired inverse (default)
        _ext.check(lib.mms_set_option(b"RESAMPLE_PAIRED", 1))
        _ext.check(lib.mms_set_option(b"RESAMPLE_PAIRED_C", 0))
        try:
            single_c = pp.resample_on_device(xd, num).cpu().numpy()
        finally:
            _ext.check(lib.mms_clear_option(b"RESAMPLE_PAIRED_C"))
        err_ok(single_c, ref)
        assert np.abs(outs[1] - single_c).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    finally:
        _ext.check(lib.mms_set_option(b"RESAMPLE_PAIRED", prev) if prev >= 0 else lib.mms_clear_option(b"RESAMPLE_PAIRED"))


@pytest.mark.parametrize("n,num,n_sig", [(70037, 6403, 3), (24000, 384000, 2), (1000, 500, 1), (190001, 17371, 2), (131, 977, 1)])
def test_resample_fast_passes_equal_generic_passes(n, num, n_sig):
    """csrc/fft_fast.cuh (register-resident mixed-radix passes, smooth lengths, fused chirp / filter / pruning: the default) against
    the round-1 power-of-two radix-2 passes through shared memory (MMS_RESAMPLE_FAST=0) and against the numpy restatement."""
    from multimodalsignal_b200 import _ext, preprocess as pp
    lib = _ext.lib()
    rng = np.random.default_rng(n + num)
    x = rng.standard_normal((n_sig, n)) * 3.0 + 1.5
    xd = torch.from_numpy(x).cuda()
    ref = np.stack([po.fft_resample(r, num) for r in x])
    prev = lib.mms_get_option(b"RESAMPLE_FAST", -1)
    try:
        outs = {}
        for mode in (1, 0):
            _ext.check(lib.mms_set_option(b"RESAMPLE_FAST", mode))
            outs[mode] = pp.resample_on_device(xd, num).cpu().numpy()
            err_ok(outs[mode], ref)
        assert np.abs(outs[1] - outs[0]).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    finally:
        _ext.check(lib.mms_set_option(b"RESAMPLE_FAST", prev) if prev >= 0 else lib.mms_clear_option(b"RESAMPLE_FAST"))


def test_resample_full_size_recording():
    """BASELINE-size stream: 700 Hz x 100 min + odd offset (N = 4 200 959 has a large prime factor)."""
    from multimodalsignal_b200 import preprocess as pp
    n = 4200000 + 137 * 7
    num = pp.resampled_length(n, 700, 64)
    rng = np.random.default_rng(5)
    t = np.arange(n) / 700.0
    x = np.stack([np.sin(2 * np.pi * 1.2 * t) + 0.3 * rng.standard_normal(n), 33.0 + 0.1 * rng.standard_normal(n)])
    xd = torch.from_numpy(x).cuda()
    got = pp.resample_on_device(xd, num)
    ref = np.stack([po.fft_resample(r, num) for r in x])
    err_ok(got.cpu().numpy(), ref, tol=1e-9)
    # size-independent properties: linearity, and a constant stays a constant
    a, b = 0.75, -1.5
    mix = pp.resample_on_device((a * xd[0] + b * xd[1])[None], num)[0]
    lin = a * got[0] + b * got[1]
    assert (mix - lin).abs().max().item() <= 1e-9 * 40
    const = pp.resample_on_device(torch.full((1, n), 2.5, dtype=torch.float64, device="cuda"), num)
    assert (const - 2.5).abs().max().item() <= 1e-10


def _short_tree(tmp_path, sids):
    root = tmp_path / "WESAD"
    synth.write_wesad_tree(root, subjects=sids, minutes=synth.SHORT_MINUTES, protocol=synth.SHORT_PROTOCOL, with_wrist=True)
    return root


@pytest.mark.parametrize("fs", [64, 128])
def test_run_preprocessing_matches_reference(tmp_path, fs, monkeypatch):
    """Files written by the drop-in run_preprocessing() vs the reference's (golden)."""
    from multimodalsignal_b200 import preprocess as pp
    z, _ = load_golden("preprocess_golden.npz")
    root = _short_tree(tmp_path, ["S2", "S5"])
    monkeypatch.setattr(pp, "RAW_FS", fs)
    done = pp.run_preprocessing(wesad_root=root, output_path=tmp_path / "data")
    assert done == ["S2", "S5"]
    out = tmp_path / "data" / "chest_raw"
    assert (out / "_channel_names.txt").read_text() == bytes(z["channel_names"]).decode()
    for sid in ("S2", "S5"):
        X, y = np.load(out / f"{sid}_X.npy"), np.load(out / f"{sid}_y.npy")
        assert X.dtype == np.float64 and y.dtype == np.int64
        assert tuple(z[f"short/{fs}/{sid}/shape"]) == X.shape
        assert np.array_equal(y, z[f"short/{fs}/{sid}/labels"])
        np.testing.assert_allclose(X[:, ::61, :], z[f"short/{fs}/{sid}/X_sub"], atol=1e-9)
        np.testing.assert_allclose(X.sum(axis=1), z[f"short/{fs}/{sid}/X_rowsum"], atol=1e-7)


def test_window_gather_is_bit_exact_copy_of_streams(tmp_path):
    from multimodalsignal_b200 import preprocess as pp
    root = _short_tree(tmp_path, ["S5"])
    data = pp.load_pkl("S5", root)
    sub = pp.preprocess_subject("S5", data, pp.parse_quest_csv("S5", root), 64, include_wrist=True)
    assert sub.streams.shape[0] == 14 and sub.channel_names[8:] == pp.WRIST_CHANNEL_NAMES
    X = sub.windows_f64().cpu().numpy()
    host = sub.streams.cpu().numpy()
    ref = np.stack([host[:, s:s + sub.window].T for s in sub.starts_host])
    assert np.array_equal(X, ref)
    # wrist extension = the same resampler applied at 32 / 64 / 4 Hz: check one channel
    wrist = {k.decode(): v for k, v in data[b'signal'][b'wrist'].items()}
    eda = po.resample_signal(wrist["EDA"][:, 0], 4, 64)
    np.testing.assert_allclose(host[12, :len(eda)], eda, atol=1e-10)


@pytest.mark.parametrize("mode", ["stress_binary", "ternary"])
def test_device_dataset_matches_reference_dataset(tmp_path, mode):
    """DeviceWesadDataset (streams -> stats -> normalised float32 [N,C,W] on the GPU) vs the reference
    WesadDataset on the reference's own npy files (golden)."""
    from multimodalsignal_b200 import preprocess as pp
    from multimodalsignal_b200.dataset import DeviceBatchLoader, DeviceWesadDataset
    z, _ = load_golden("preprocess_golden.npz")
    root = _short_tree(tmp_path, ["S2", "S5"])
    streams = {sid: pp.preprocess_subject(sid, pp.load_pkl(sid, root), pp.parse_quest_csv(sid, root), 64) for sid in ("S2", "S5")}
    chans = ["chest_ECG", "chest_EDA", "chest_EMG", "chest_Resp"]
    ds = DeviceWesadDataset(streams, ["S2", "S5", "S99"], chans, classification_mode=mode)
    assert len(ds) == int(z[f"dataset/{mode}/len"])
    assert np.array_equal(ds.labels.cpu().numpy(), z[f"dataset/{mode}/labels"])
    data = ds.data.cpu().numpy()                                   # [N, C, W] float32
    np.testing.assert_allclose(data[:, :, ::61].transpose(0, 2, 1), z[f"dataset/{mode}/data_sub"], atol=2e-5)
    x3, y3 = ds[3]
    assert x3.dtype == torch.float32 and tuple(x3.shape) == (4, 3840)
    np.testing.assert_allclose(x3.cpu().numpy()[:, ::61], z[f"dataset/{mode}/item3_x_sub"], atol=2e-5)
    assert int(y3.item()) == int(z[f"dataset/{mode}/item3_y"])
    seen = 0
    for xb, yb in DeviceBatchLoader(ds, 8, shuffle=True):
        assert xb.is_cuda and xb.shape[1:] == (4, 3840) and yb.dtype == torch.int64
        seen += xb.shape[0]
    assert seen == len(ds)
