"""Drop-in for the reference ``dataset.py`` plus the device-resident dataset.

``WesadDataset`` keeps the reference contract exactly: reads ``{sid}_X.npy / {sid}_y.npy``,
selects channels, maps labels, normalises per subject over the WINDOWED array (``chest_EDA``
through ``log1p`` with its own statistics, dataset.py:37-48), concatenates; ``__getitem__`` yields
``(float32 [C, W], int64 scalar)``.  This file-based path is host-side data plumbing (numpy), as in
the reference; it feeds ``DataLoader`` -> the CUDA training step.

``DeviceWesadDataset`` is the B200-native replacement for the same contract: it is built from the
continuous resampled streams (``preprocess.SubjectStreams``), computes the same overlap-weighted
statistics with ``mms_window_stats`` and materialises the normalised, already-permuted float32
``[N, C, W]`` tensor once on the device with ``mms_window_gather`` -- no 6x-expanded float64 window
array, no per-item float64->float32 conversion, no per-step H2D copy.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _ext
from ._ext import check, ptr, stream


def map_labels(y_raw: np.ndarray, classification_mode: str) -> np.ndarray:
    """reference dataset.py:28-34."""
    if classification_mode == 'stress_binary':
        return np.where(y_raw == 2, 1, 0)
    if classification_mode == 'ternary':
        return np.where(y_raw == 1, 0, np.where(y_raw == 3, 1, np.where(y_raw == 2, 2, 0)))
    raise ValueError(f"Unknown classification_mode: {classification_mode}")


class WesadDataset(Dataset):
    """reference dataset.py:8-65 (same constructor, attributes ``data`` / ``labels``, item shapes)."""

    def __init__(self, data_path: Path, subjects: list, channels_to_use: list,
                 all_channel_names: list, classification_mode='stress_binary'):
        self.classification_mode = classification_mode
        data_path = Path(data_path)
        channel_indices = [all_channel_names.index(ch) for ch in channels_to_use]
        data_list, labels_list = [], []
        for sid in subjects:
            x_file, y_file = data_path / f'{sid}_X.npy', data_path / f'{sid}_y.npy'
            if not x_file.exists() or not y_file.exists():
                print(f"Warning: Skipping subject {sid} for data, file not found.")
                continue
            x_sel = np.load(x_file)[:, :, channel_indices]
            y = map_labels(np.load(y_file), classification_mode)
            mean_all = np.mean(x_sel, axis=(0, 1))
            std_all = np.std(x_sel, axis=(0, 1)) + 1e-8
            for ch, name in enumerate(channels_to_use):
                if name == 'chest_EDA':
                    log_data = np.log1p(x_sel[:, :, ch])
                    x_sel[:, :, ch] = (log_data - np.mean(log_data)) / (np.std(log_data) + 1e-8)
                else:
                    x_sel[:, :, ch] = (x_sel[:, :, ch] - mean_all[ch]) / std_all[ch]
            data_list.append(x_sel)
            labels_list.append(y)
        if not data_list:
            raise ValueError(f"No data loaded for subjects: {subjects}. Check paths and data existence.")
        self.data = np.concatenate(data_list, axis=0)
        self.labels = np.concatenate(labels_list, axis=0)

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, idx):
        x = torch.from_numpy(self.data[idx]).float().permute(1, 0)
        y = torch.tensor(self.labels[idx], dtype=torch.long)
        return x, y


def normalise_gather(sub, channels_to_use):
    """One subject: overlap-weighted statistics + normalised float32 ``[n_win, C, W]`` on the device
    (the arithmetic of dataset.py:37-48 followed by dataset.py:63's cast and permute)."""
    cache = sub.__dict__.setdefault("_normalised", {})      # the statistics are per subject: every LOSO fold reuses them
    key = tuple(channels_to_use)
    if key in cache:
        return cache[key]
    lib = _ext.lib()
    idx = [sub.channel_names.index(ch) for ch in channels_to_use]
    rows = [sub.streams[i] for i in idx]
    arr = (C.c_void_p * len(rows))(*[r.data_ptr() for r in rows])
    dev = sub.streams.device
    n_win, W = len(sub.labels), sub.window
    flags = torch.tensor([1 if ch == 'chest_EDA' else 0 for ch in channels_to_use], dtype=torch.int32, device=dev)
    sums = torch.zeros(len(idx), 2, dtype=torch.float64, device=dev)
    check(lib.mms_window_stats(arr, len(idx), sub.streams.shape[1], ptr(sub.starts), n_win, W, ptr(flags), ptr(sums), stream()))
    cnt = float(n_win) * float(W)
    mean = sums[:, 0] / cnt
    var = (sums[:, 1] / cnt - mean * mean).clamp_min(0.0)
    scale = 1.0 / (var.sqrt() + 1e-8)                                   # dataset.py:38,45: std + 1e-8
    out = torch.empty(n_win, len(idx), W, dtype=torch.float32, device=dev)
    check(lib.mms_window_gather(arr, len(idx), sub.streams.shape[1], ptr(sub.starts), n_win, W, 1,
                                ptr(mean.contiguous()), ptr(scale.contiguous()), ptr(flags), ptr(out), stream()))
    cache[key] = out
    return out


class DeviceWesadDataset(Dataset):
    """Same contract as ``WesadDataset`` but device-resident: ``data`` is float32 CUDA ``[N, C, W]``
    (already in ``__getitem__`` layout), ``labels`` int64 CUDA ``[N]``.  ``subject_streams`` maps
    subject id -> ``preprocess.SubjectStreams``; missing subjects are skipped with the reference's
    warning (dataset.py:20-22)."""

    def __init__(self, subject_streams: dict, subjects: list, channels_to_use: list,
                 classification_mode='stress_binary'):
        self.classification_mode = classification_mode
        xs, ys = [], []
        for sid in subjects:
            sub = subject_streams.get(sid)
            if sub is None or len(sub.labels) == 0:
                print(f"Warning: Skipping subject {sid} for data, file not found.")
                continue
            xs.append(normalise_gather(sub, channels_to_use))
            ys.append(torch.from_numpy(map_labels(sub.labels, classification_mode)).to(sub.streams.device))
        if not xs:
            raise ValueError(f"No data loaded for subjects: {subjects}. Check paths and data existence.")
        self.data = torch.cat(xs, dim=0)
        self.labels = torch.cat(ys, dim=0).long()

    def __len__(self):
        return int(self.labels.shape[0])

    def __getitem__(self, idx):
        return self.data[idx], self.labels[idx]


class DeviceBatchLoader:
    """``DataLoader(dataset, batch_size, shuffle)`` for a ``DeviceWesadDataset``: batches are
    index-gathers on the device (``torch.randperm`` on the CPU generator keeps the reference's
    shuffling semantics: a fresh permutation per epoch)."""

    def __init__(self, dataset: DeviceWesadDataset, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 generator=None):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.generator = generator          # as DataLoader(generator=...): the CPU generator the permutations come from

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        order = order.to(self.dataset.data.device)
        for i in range(len(self)):
            sel = order[i * self.batch_size:(i + 1) * self.batch_size]
            yield self.dataset.data.index_select(0, sel), self.dataset.labels.index_select(0, sel)
