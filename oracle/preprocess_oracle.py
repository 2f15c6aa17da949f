"""CPU restatement (numpy) of the reference resample + windowing + normalisation.

TEST INFRASTRUCTURE ONLY -- see ``oracle/model_oracle.py`` for the import rule.

The resampler the reference calls lives in a third-party dependency that is not
vendored under ``/root/reference``: ``scipy.signal.resample`` (reference
requirements.txt:7 pins scipy 1.15.3 by comment; this container has 1.18.1).  Its
published algorithm for real input and no window (``scipy/signal/_signaltools.py``)
is restated in :func:`fft_resample`.  Pinning: ``tests/test_oracle_preprocess.py``
checks this file against ``scipy.signal.resample`` itself and against
``tests/golden/preprocess_*.npz`` produced by running the UNMODIFIED reference
``preprocess.py`` / ``dataset.py`` on the synthetic recordings
(``oracle/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

ORIGINAL_CHEST_FS = 700                      # reference preprocess.py:17
CHEST_CHANNELS = ["ACC", "ECG", "EDA", "EMG", "Resp", "Temp"]   # preprocess.py:27
TASK_TO_LABEL_MAP = {"Base": 1, "TSST": 2, "Fun": 3, "Medi1": 4, "Medi2": 4}  # preprocess.py:28
WINDOW_SEC, STRIDE_SEC = 60, 10              # preprocess.py:22-23


def resampled_length(n: int, original_fs, target_fs) -> int:
    """reference preprocess.py:72,74 -- ``int(len * (target_fs / original_fs))`` in
    float64, in exactly this operation order."""
    return int(n * (target_fs / original_fs))


def fft_resample(x: np.ndarray, num: int) -> np.ndarray:
    """``scipy.signal.resample(x, num)`` for 1-D real ``x`` (Fourier method):
    rfft, keep the ``min(num, N)//2 + 1`` lowest bins, fix the unpaired bin when the
    kept count is even, inverse real FFT of length ``num`` scaled by ``num / N``."""
    x = np.asarray(x, dtype=np.float64)
    n_x = x.shape[0]
    X = np.fft.rfft(x)
    m = min(num, n_x)
    m2 = m // 2 + 1
    X = X[:m2].copy()
    if m % 2 == 0 and num != n_x:
        X[m // 2] *= 2.0 if num < n_x else 0.5
    return np.fft.irfft(X / (n_x / num), n=num)


def resample_signal(signal_data: np.ndarray, original_fs, target_fs) -> np.ndarray:
    """reference preprocess.py:70-75."""
    num = resampled_length(len(signal_data), original_fs, target_fs)
    if signal_data.ndim > 1:
        return np.column_stack([fft_resample(signal_data[:, i], num) for i in range(signal_data.shape[1])])
    return fft_resample(signal_data, num)


def apply_subject_quirk(sid: str, protocol):
    """reference preprocess.py:53-57 -- for S2 and S6 the Base segment starts at the
    midpoint of its original span."""
    protocol = [list(p) for p in protocol]
    if sid in ("S2", "S6"):
        for row in protocol:
            if row[0] == "Base":
                row[1] = (row[1] + row[2]) / 2
                break
    return [tuple(p) for p in protocol]


def window_plan(protocol, target_fs, original_fs=ORIGINAL_CHEST_FS,
                window_sec=WINDOW_SEC, stride_sec=STRIDE_SEC):
    """Window start indices (in resampled samples) and raw labels.

    reference preprocess.py:160-167 and 185-189; every product is evaluated
    left-to-right in float64 and truncated with ``int`` exactly as the reference
    does -- never "fixed" to exact arithmetic (SURVEY §8d float traps)."""
    starts, labels = [], []
    w = int(window_sec * target_fs)
    stride = int(stride_sec * target_fs)
    for task, start_min, end_min in protocol:
        label = TASK_TO_LABEL_MAP.get(task.replace(" ", "").strip())
        if label is None:
            continue
        start_idx_orig = int(start_min * 60 * original_fs)
        end_idx_orig = int(end_min * 60 * original_fs)
        start_idx = int(start_idx_orig * (target_fs / original_fs))
        end_idx = int(end_idx_orig * (target_fs / original_fs))
        for i in range(start_idx, end_idx - w + 1, stride):
            starts.append(i)
            labels.append(label)
    return np.asarray(starts, dtype=np.int64), np.asarray(labels, dtype=np.int64), w


def stack_windows(resampled: dict, starts: np.ndarray, w: int) -> np.ndarray:
    """reference preprocess.py:189-200, 218 -- per window, slice every sensor and
    concatenate columns in CHEST_CHANNELS order -> ``[N_win, W, 8]`` float64.
    Slices that run past the end of a stream are ragged in the reference (numpy
    slicing clips); the synthetic protocols never do that and this oracle raises."""
    cols = np.concatenate([resampled[c].reshape(len(resampled[c]), -1) for c in CHEST_CHANNELS], axis=1)
    if len(starts) and starts.max() + w > cols.shape[0]:
        raise ValueError("window runs past the end of the resampled stream")
    return np.stack([cols[s:s + w] for s in starts]) if len(starts) else np.zeros((0, w, cols.shape[1]))


def preprocess_subject(sid: str, chest: dict, protocol, target_fs):
    """One iteration of the reference subject loop (preprocess.py:138-222) for the
    'raw' target.  Returns ``(X [N,W,8] f64, y [N] i64)``."""
    resampled = {c: resample_signal(chest[c], ORIGINAL_CHEST_FS, target_fs) for c in CHEST_CHANNELS}
    starts, labels, w = window_plan(apply_subject_quirk(sid, protocol), target_fs)
    return stack_windows(resampled, starts, w), labels


# --------------------------------------------------------------------------- a17
def map_labels(y_raw: np.ndarray, mode: str) -> np.ndarray:
    """reference dataset.py:28-34."""
    if mode == "stress_binary":
        return np.where(y_raw == 2, 1, 0)
    if mode == "ternary":
        return np.where(y_raw == 1, 0, np.where(y_raw == 3, 1, np.where(y_raw == 2, 2, 0)))
    raise ValueError(f"Unknown classification_mode: {mode}")


def normalise_subject(x_selected: np.ndarray, channel_names: list) -> np.ndarray:
    """reference dataset.py:37-48 -- per-subject z-score over the WINDOWED array
    (overlap-weighted), ``chest_EDA`` goes through log1p with its own statistics."""
    x = np.array(x_selected, dtype=np.float64, copy=True)
    mean_all = np.mean(x, axis=(0, 1))
    std_all = np.std(x, axis=(0, 1)) + 1e-8
    for ch, name in enumerate(channel_names):
        if name == "chest_EDA":
            log_data = np.log1p(x[:, :, ch])
            x[:, :, ch] = (log_data - np.mean(log_data)) / (np.std(log_data) + 1e-8)
        else:
            x[:, :, ch] = (x[:, :, ch] - mean_all[ch]) / std_all[ch]
    return x
