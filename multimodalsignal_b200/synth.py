"""Synthetic WESAD-shaped recordings.

There is no dataset in this environment, so every parity test, the bench and the
CPU baseline run on the same seeded synthetic recordings.  The on-disk layout is
exactly what the reference loaders read:

* ``WESAD/{sid}/{sid}.pkl``  -- ``{b'signal': {b'chest': {...}, b'wrist': {...}}}`` with
  bytes keys (reference preprocess.py:64 loads with ``encoding='bytes'`` and
  preprocess.py:143-144 decodes the chest keys).
* ``WESAD/{sid}/{sid}_quest.csv`` -- ``;`` separated ``# ORDER`` / ``# START`` / ``# END``
  rows with an equal field count on every line (reference preprocess.py:43-49).

The protocol start times 32.05, 48.07 and 70.13 minutes are chosen because
``int(start_min * 60 * 700)`` evaluated in float64 differs from exact arithmetic for
them (the "float trap" of reference preprocess.py:166-167); index parity tests
depend on that.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ALL_SUBJECTS = [f"S{i}" for i in range(2, 18) if i != 12]  # reference main.py:67

CHEST_FS = 700
WRIST_FS = {"ACC": 32, "BVP": 64, "EDA": 4, "TEMP": 4}

# (task, start_min, end_min); 'sRead' is not in the reference TASK_TO_LABEL_MAP
# (preprocess.py:28) and must be skipped by the windowing code.
FULL_PROTOCOL = [
    ("Base", 5.0, 25.0),
    ("TSST", 32.05, 43.5),
    ("Medi 1", 48.07, 55.0),
    ("Fun", 58.5, 65.0),
    ("Medi 2", 70.13, 77.0),
    ("sRead", 78.0, 79.5),
]
FULL_MINUTES = 100.0  # 700 * 6000 = 4 200 000 chest samples

# A short protocol for CPU-sized tests: every segment still holds at least one
# 60 s window and the float-trap start times are kept.
SHORT_PROTOCOL = [
    ("Base", 0.5, 3.0),
    ("TSST", 3.07, 4.5),
    ("Medi 1", 4.56, 5.8),
    ("Fun", 6.0, 7.2),
    ("Medi 2", 7.3, 8.5),
    ("sRead", 8.6, 8.9),
]
SHORT_MINUTES = 9.0

_TASK_AMPLITUDE = {"Base": 1.0, "TSST": 1.8, "Medi 1": 0.8, "Fun": 1.3, "Medi 2": 0.8, "sRead": 1.0}


@dataclass
class SyntheticSubject:
    sid: str
    chest: dict                      # name -> float64 [N, k]
    wrist: dict                      # name -> float64 [n, k]
    protocol: list = field(default_factory=list)

    def as_pickle_dict(self) -> dict:
        enc = lambda d: {k.encode(): v for k, v in d.items()}
        return {b"signal": {b"chest": enc(self.chest), b"wrist": enc(self.wrist)},
                b"subject": self.sid.encode()}


def _amplitude_track(n: int, fs: float, protocol) -> np.ndarray:
    amp = np.ones(n)
    for task, s, e in protocol:
        a, b = int(s * 60 * fs), min(n, int(e * 60 * fs))
        if a < b:
            amp[a:b] = _TASK_AMPLITUDE.get(task, 1.0)
    return amp


def _oscillation(rng, n: int, fs: float, k: int, amp: np.ndarray) -> np.ndarray:
    t = np.arange(n) / fs
    cols = []
    for _ in range(k):
        ph = rng.uniform(0, 2 * np.pi, 3)
        base = (np.sin(2 * np.pi * 1.2 * t + ph[0]) + 0.6 * np.sin(2 * np.pi * 0.25 * t + ph[1])
                + 0.3 * np.sin(2 * np.pi * min(17.0, 0.4 * fs) * t + ph[2]))
        cols.append(amp * base + 0.3 * rng.standard_normal(n))
    return np.stack(cols, axis=1)


def make_subject(sid: str, idx: int, seed: int = 42, minutes: float = FULL_MINUTES,
                 protocol=None, with_wrist: bool = True) -> SyntheticSubject:
    """One synthetic subject.  ``idx`` adds an odd offset to the chest length so the
    FFT lengths are not smooth (real recordings have arbitrary lengths)."""
    protocol = FULL_PROTOCOL if protocol is None else protocol
    rng = np.random.default_rng([seed, idx])
    n = int(CHEST_FS * 60 * minutes) + 137 * idx
    amp = _amplitude_track(n, CHEST_FS, protocol)
    t = np.arange(n) / CHEST_FS
    chest = {
        "ACC": _oscillation(rng, n, CHEST_FS, 3, amp),
        "ECG": _oscillation(rng, n, CHEST_FS, 1, amp),
        "EDA": (2.0 + 0.5 * np.sin(2 * np.pi * 0.01 * t) * amp
                + np.abs(rng.normal(0, 0.05, n)))[:, None],          # strictly > 0 (log1p, dataset.py:43)
        "EMG": _oscillation(rng, n, CHEST_FS, 1, amp),
        "Resp": _oscillation(rng, n, CHEST_FS, 1, amp),
        "Temp": (33.0 + rng.normal(0, 0.1, n))[:, None],
    }
    wrist = {}
    if with_wrist:
        for name, fs in WRIST_FS.items():
            nw = int(fs * 60 * minutes) + (idx if fs >= 32 else 0)
            wamp = _amplitude_track(nw, fs, protocol)
            tw = np.arange(nw) / fs
            if name == "ACC":
                wrist[name] = _oscillation(rng, nw, fs, 3, wamp)
            elif name == "BVP":
                wrist[name] = _oscillation(rng, nw, fs, 1, wamp)
            elif name == "EDA":
                wrist[name] = (1.5 + 0.4 * np.sin(2 * np.pi * 0.01 * tw) * wamp
                               + np.abs(rng.normal(0, 0.05, nw)))[:, None]
            else:
                wrist[name] = (32.0 + rng.normal(0, 0.1, nw))[:, None]
    return SyntheticSubject(sid, chest, wrist, list(protocol))


def quest_csv_text(protocol) -> str:
    """``{sid}_quest.csv`` body with the same number of fields on every line."""
    tasks = [p[0] for p in protocol]
    width = len(tasks) + 1
    pad = lambda cells: ";".join(cells + [""] * (width - len(cells)))
    lines = [
        pad(["# Subj"]),
        pad(["# ORDER"] + tasks),
        pad(["# START"] + [repr(float(p[1])) for p in protocol]),
        pad(["# END"] + [repr(float(p[2])) for p in protocol]),
    ]
    return "\n".join(lines) + "\n"


def write_wesad_tree(root: Path, subjects=None, seed: int = 42, minutes: float = FULL_MINUTES,
                     protocol=None, with_wrist: bool = True) -> list:
    """Write ``root/{sid}/{sid}.pkl`` and ``{sid}_quest.csv`` for each subject."""
    root = Path(root)
    subjects = ALL_SUBJECTS if subjects is None else subjects
    out = []
    for sid in subjects:
        idx = ALL_SUBJECTS.index(sid) if sid in ALL_SUBJECTS else len(out)
        sub = make_subject(sid, idx, seed=seed, minutes=minutes, protocol=protocol, with_wrist=with_wrist)
        d = root / sid
        d.mkdir(parents=True, exist_ok=True)
        with open(d / f"{sid}.pkl", "wb") as f:
            pickle.dump(sub.as_pickle_dict(), f, protocol=4)
        (d / f"{sid}_quest.csv").write_text(quest_csv_text(sub.protocol))
        out.append(sub)
    return out


def synthetic_windows(n: int, channels: int, length: int, seed: int = 0, num_classes: int = 2):
    """Random ``float32 [n, channels, length]`` windows + ``int64 [n]`` labels of the
    shape ``WesadDataset.__getitem__`` yields (reference dataset.py:62-65)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, channels, length)).astype(np.float32)
    y = rng.integers(0, num_classes, size=n).astype(np.int64)
    # make the label weakly recoverable so accuracy parity is not trivial
    x[:, 0, :] += (y[:, None] * 0.5).astype(np.float32)
    return x, y
