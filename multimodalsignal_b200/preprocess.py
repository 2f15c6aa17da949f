"""Drop-in for the resample + windowing part of the reference ``preprocess.py`` ('raw' target).

Same module constants, ``resample_signal(signal_data, original_fs, target_fs)``,
``parse_quest_csv``, ``load_pkl`` and ``run_preprocessing()``; same output files
(``data/chest_raw/{sid}_X.npy [N_win, W, 8] float64``, ``{sid}_y.npy [N_win] int64``,
``_channel_names.txt``).  The FFT resampling (scipy.signal.resample semantics) and the window
stacking run as sm_100a kernels (``mms_resample_f64`` / ``mms_window_gather``); the window start
indices are computed on the host in float64 with the reference's exact expressions
(preprocess.py:166-167,185-188), so they are bit-identical -- float traps included.

Differences from the reference as shipped, all deliberate and documented:
  * ``RAW_FS`` defaults to 64 (README / north star) instead of 128; it stays a module constant;
  * ``PROCESS_TARGETS`` defaults to ``['raw']`` -- the directory ``main.py`` actually reads (D6);
    the 'feature' / 'raw-align' branches need neurokit2 and are out of scope;
  * wrist streams (extension, D2): ``include_wrist=True`` also resamples wrist ACC/BVP/EDA/TEMP
    (32/64/4/4 Hz) with the same resampler and writes ``data/all_raw`` with 14 channels.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import threading
from pathlib import Path

import numpy as np
import torch

from . import _ext
from ._ext import check, ptr, stream

# --- reference module constants (preprocess.py:13-28) -------------------------------------------
WESAD_ROOT = Path('./WESAD')
OUTPUT_PATH = Path('./data')
ORIGINAL_CHEST_FS = 700
PROCESS_TARGETS = ['raw']
RAW_FS = 64
RAW_WINDOW_SEC = 60
RAW_STRIDE_SEC = 10
CHEST_CHANNELS = ['ACC', 'ECG', 'EDA', 'EMG', 'Resp', 'Temp']
TASK_TO_LABEL_MAP = {'Base': 1, 'TSST': 2, 'Fun': 3, 'Medi1': 4, 'Medi2': 4}
WRIST_CHANNELS = {'ACC': 32, 'BVP': 64, 'EDA': 4, 'TEMP': 4}           # extension

CHEST_CHANNEL_NAMES = [f"chest_ACC_{ax}" for ax in 'xyz'] + [f"chest_{c}" for c in ['ECG', 'EDA', 'EMG', 'Resp', 'Temp']]
WRIST_CHANNEL_NAMES = [f"wrist_ACC_{ax}" for ax in 'xyz'] + ['wrist_BVP', 'wrist_EDA', 'wrist_TEMP']

_MAX_SIGNALS_PER_CALL = 8


def resampled_length(n: int, original_fs, target_fs) -> int:
    """``int(len * (target_fs / original_fs))`` -- reference preprocess.py:72, same float64 order."""
    return int(n * (target_fs / original_fs))


def resample_on_device(x: torch.Tensor, num: int) -> torch.Tensor:
    """``x`` float64 CUDA ``[n_sig, N]`` (rows contiguous) -> float64 ``[n_sig, num]``."""
    lib = _ext.lib()
    if x.dtype != torch.float64 or not x.is_cuda or x.dim() != 2:
        raise _ext.MmsError("resample_on_device expects a float64 CUDA tensor [n_sig, N]")
    x = x.contiguous()
    n_sig, n = x.shape
    y = torch.empty(n_sig, num, dtype=torch.float64, device=x.device)
    for s0 in range(0, n_sig, _MAX_SIGNALS_PER_CALL):
        k = min(_MAX_SIGNALS_PER_CALL, n_sig - s0)
        nbytes = lib.mms_resample_workspace_bytes(n, num, k)
        if nbytes < 0:
            raise _ext.MmsError(f"resample: unsupported lengths {n} -> {num}")
        ws = _workspace(x.device, int(nbytes))
        check(lib.mms_resample_f64(ptr(x[s0:s0 + k]), n, num, k, ptr(y[s0:s0 + k]), ptr(ws), int(nbytes), stream()))
    return y


_WORKSPACES = {}


def _workspace(device, nbytes):
    """One grow-only resample workspace per (device, stream): calls on a stream run in order, so they can share it, and the
    caching allocator is not asked for a fresh 0.1-0.5 GB block per call and stream (every recording has its own length, so
    those requests rarely hit a cached block of the right size; an occasional pass over the subjects ran 2.5x slower)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


_SIDE_STREAMS = {}
_MANY_ORDER = int(os.environ.get("MMS_RESAMPLE_MANY_ORDER", "1"))     # 0 = everything on the current stream (A/B)


def resample_many(items):
    """``items``: list of ``(x [k, N] float64 CUDA, num)``.  The first (large) one runs on the current stream, every other one
    on its own HIGH-PRIORITY side stream beside it.  The wrist channels of a subject are short transforms of ~20 launches
    each whose grids do not fill the GPU: a launch lasts as long as one CTA (10-25 us), so run one after the other they cost
    about as much device time as the 30x larger chest call.  Beside each other their latencies overlap, and with priority
    their few CTAs get the slots the chest kernels' CTAs free every couple of microseconds instead of waiting for a kernel
    boundary."""
    if len(items) == 1 or _MANY_ORDER == 0:
        return [resample_on_device(x, n) for x, n in items]
    dev = items[0][0].device
    key = (dev.index, threading.get_ident())
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < len(items) - 1:
        pool.append(torch.cuda.Stream(device=dev, priority=-1))
    cur = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(cur)                       # the inputs exist here
    rest = []
    for (x, n), side in zip(items[1:], pool):
        side.wait_event(ready)
        with torch.cuda.stream(side):
            rest.append(resample_on_device(x, n))
    first = resample_on_device(*items[0])
    for y, side in zip(rest, pool):
        cur.wait_stream(side)
        y.record_stream(cur)
    return [first] + rest


def resample_subject_rows(chest_rows, wrist_rows=None, target_fs=None):
    """Chest rows ``[8, N]`` at 700 Hz (+ the wrist rows of ``WRIST_CHANNELS``, each at its own rate) -> float64
    ``[n_channels, num]`` at ``target_fs``; wrist streams that end a few samples early are padded with their last value
    (the wrist clock is not the chest clock), longer ones cut.  Wrist channels of equal rate and length share a call."""
    target_fs = RAW_FS if target_fs is None else target_fs
    num = resampled_length(chest_rows.shape[1], ORIGINAL_CHEST_FS, target_fs)
    items = [(chest_rows, num)]
    if wrist_rows:
        groups = []                                     # [fs, n, [rows...]] of consecutive channels (the output keeps their order)
        for name, fs in WRIST_CHANNELS.items():
            r = wrist_rows[name]
            if groups and groups[-1][0] == fs and groups[-1][1] == r.shape[1]:
                groups[-1][2].append(r)
            else:
                groups.append([fs, r.shape[1], [r]])
        for fs, n, rs in groups:
            items.append((rs[0] if len(rs) == 1 else torch.cat(rs, dim=0), resampled_length(n, fs, target_fs)))
    ys = resample_many(items)
    if len(ys) == 1:
        return ys[0]
    out = torch.empty(sum(y.shape[0] for y in ys), num, dtype=torch.float64, device=chest_rows.device)
    r0 = 0
    for y in ys:
        k, nw = y.shape
        out[r0:r0 + k, :min(nw, num)] = y[:, :num]
        if nw < num:
            out[r0:r0 + k, nw:] = y[:, -1:]
        r0 += k
    return out


def resample_signal(signal_data, original_fs, target_fs):
    """reference preprocess.py:70-75.  ``signal_data``: real ndarray ``[N]`` or ``[N, k]`` ->
    float64 ndarray ``[num]`` / ``[num, k]`` with ``num = int(N * (target_fs / original_fs))``."""
    arr = np.asarray(signal_data)
    num = resampled_length(len(arr), original_fs, target_fs)
    cols = arr.reshape(len(arr), -1).T                                   # [k, N]
    x = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.float64)).cuda()
    y = resample_on_device(x, num).cpu().numpy()
    return np.ascontiguousarray(y.T) if arr.ndim > 1 else y[0]


def parse_quest_csv(subject_id: str, wesad_root: Path):
    """reference preprocess.py:41-58 -> list of ``(task, start_min, end_min)`` (the reference
    returns the same three columns as a DataFrame).  Includes the S2 / S6 quirk: the Base segment
    starts at the midpoint of its span."""
    quest_path = Path(wesad_root) / subject_id / f"{subject_id}_quest.csv"
    rows = {}
    for line in quest_path.read_text().splitlines():
        if not line.strip():
            continue
        cells = line.split(';')
        for tag in ('# ORDER', '# START', '# END'):
            if tag in cells[0] and tag not in rows:
                rows[tag] = [c.strip() for c in cells[1:] if c.strip() != '']
    if len(rows) != 3:
        raise ValueError(f"{quest_path}: missing # ORDER / # START / # END rows")
    tasks = rows['# ORDER']
    starts = [float(v) for v in rows['# START']]
    ends = [float(v) for v in rows['# END']]
    if not (len(tasks) == len(starts) == len(ends)):
        raise ValueError(f"为受试者 {subject_id} 解析出的任务、开始、结束时间长度不匹配!")
    return base_halving_quirk(subject_id, zip(tasks, starts, ends))


def base_halving_quirk(subject_id: str, protocol):
    """reference preprocess.py:53-57: the Base segment of S2 and S6 starts at the midpoint of its span."""
    protocol = [list(row) for row in protocol]
    if subject_id in ['S2', 'S6']:
        for row in protocol:
            if row[0] == 'Base':
                row[1] = (row[1] + row[2]) / 2
                break
    return [tuple(r) for r in protocol]


def load_pkl(subject_id: str, wesad_root: Path):
    pkl_path = Path(wesad_root) / subject_id / f"{subject_id}.pkl"
    try:
        with open(pkl_path, 'rb') as f:
            return pickle.load(f, encoding='bytes')
    except FileNotFoundError:
        print(f"警告: 无法找到文件 {pkl_path}")
        return None


def window_plan(protocol, target_fs, original_fs=ORIGINAL_CHEST_FS,
                window_sec=RAW_WINDOW_SEC, stride_sec=RAW_STRIDE_SEC):
    """Window start indices / raw labels / window length, reference preprocess.py:160-167,185-189.
    Every product is evaluated left to right in float64 and truncated with ``int`` exactly as
    the reference does."""
    starts, labels = [], []
    window = int(window_sec * target_fs)
    stride = int(stride_sec * target_fs)
    for task, start_min, end_min in protocol:
        label = TASK_TO_LABEL_MAP.get(task.replace(" ", "").strip())
        if label is None:
            continue
        start_idx_orig = int(start_min * 60 * original_fs)
        end_idx_orig = int(end_min * 60 * original_fs)
        start_idx_raw = int(start_idx_orig * (target_fs / original_fs))
        end_idx_raw = int(end_idx_orig * (target_fs / original_fs))
        for i in range(start_idx_raw, end_idx_raw - window + 1, stride):
            starts.append(i)
            labels.append(label)
    return np.asarray(starts, dtype=np.int64), np.asarray(labels, dtype=np.int64), window


def _stream_rows(sensor_dict, names):
    """Stack the sensors of one device into float64 rows ``[n_channels, N]`` (ACC gives 3 rows)."""
    rows = []
    for name in names:
        a = np.asarray(sensor_dict[name])
        a = a.reshape(len(a), -1)
        rows.extend(np.ascontiguousarray(a[:, j], dtype=np.float64) for j in range(a.shape[1]))
    return np.stack(rows)


class _Uploader:
    """Host -> device staging for the raw recordings (4.2 M samples x 8 chest columns = 269 MB per subject).

    ``torch.from_numpy(rows).to(device)`` costs two pageable host copies plus a pageable H2D per subject -- more
    than the resampling itself.  Here every sensor column is converted / de-interleaved ONCE, in 8 MB pieces, by a small
    thread pool (numpy releases the GIL while copying) into a ring of pinned chunks; each piece goes up with its own
    asynchronous copy on a private upload stream as soon as it is staged, so the H2D copies of a subject overlap the staging
    of its later pieces and the device work of the previous subject.  The ring is small (128 MB): page-locking a buffer for
    a whole subject cost ~250 ms per buffer (measured), several times the staging itself.
    ``groups`` stages several sensors of different lengths (chest + every wrist sensor of a subject) in one go."""

    CHUNK = 1 << 20          # doubles per pinned chunk (8 MB)
    RING = 16                # 128 MB of pinned memory in all

    def __init__(self):
        self._ring = None
        self._events = [None] * self.RING
        self._pool = None
        self._stream = {}

    @staticmethod
    def _columns(sensor_dict, names):
        cols = []
        for name in names:
            a = np.asarray(sensor_dict[name])
            a = a.reshape(len(a), -1)
            cols.extend(a[:, j] for j in range(a.shape[1]))
        return cols

    def groups(self, column_groups, device):
        """``column_groups``: list of lists of equally long 1-D arrays.  Returns one float64 device tensor ``[len(group), n]``
        per group (views of ONE device buffer).  The caller's stream waits for the upload; the host does not."""
        import concurrent.futures as cf
        offs, total, tasks = [], 0, []
        for cols in column_groups:
            n = len(cols[0])
            assert all(len(c) == n for c in cols), "the columns of a group must have the same length"
            offs.append(total)
            for c, col in enumerate(cols):
                for lo in range(0, n, self.CHUNK):
                    tasks.append((total + c * n + lo, col[lo:lo + self.CHUNK]))
            total += len(cols) * n
            total += total & 1                       # keep every group 16-byte aligned
        if self._ring is None:
            self._ring = torch.empty(self.RING * self.CHUNK, dtype=torch.float64).pin_memory()
            self._pool = cf.ThreadPoolExecutor(max_workers=min(12, max(4, (os.cpu_count() or 8) - 2)), thread_name_prefix="mms-upload")
        ring_np = self._ring.numpy()
        dev = torch.empty(total, dtype=torch.float64, device=device)
        key = (device.type, device.index)
        if key not in self._stream:
            self._stream[key] = torch.cuda.Stream(device=device)
        up = self._stream[key]
        up.wait_stream(torch.cuda.current_stream(device))           # `dev` may reuse memory the compute stream is still reading
        futures = [None] * len(tasks)

        def submit(t):
            slot = t % self.RING
            if self._events[slot] is not None:
                self._events[slot].synchronize()                     # the copy that last read this chunk has finished
            _, src = tasks[t]
            futures[t] = self._pool.submit(np.copyto, ring_np[slot * self.CHUNK:slot * self.CHUNK + len(src)], src, "unsafe")

        for t in range(min(self.RING, len(tasks))):
            submit(t)
        for t, (off, src) in enumerate(tasks):
            futures[t].result()
            slot = t % self.RING
            with torch.cuda.stream(up):
                dev[off:off + len(src)].copy_(self._ring[slot * self.CHUNK:slot * self.CHUNK + len(src)], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            self._events[slot] = ev
            if t + self.RING < len(tasks):
                submit(t + self.RING)
        torch.cuda.current_stream(device).wait_stream(up)
        return [dev[off:off + len(cols) * len(cols[0])].view(len(cols), len(cols[0])) for cols, off in zip(column_groups, offs)]

    def rows(self, sensor_dict, names, device):
        """The sensors ``names`` of one device as float64 rows ``[n_channels, N]`` on ``device`` (ACC gives 3 rows)."""
        return self.groups([self._columns(sensor_dict, names)], device)[0]


_UPLOADER = _Uploader()


class SubjectStreams:
    """Resampled continuous streams of one subject on the device + its window plan.

    ``streams`` float64 CUDA ``[n_channels, num]``; ``starts`` int64 CUDA ``[n_win]``; ``labels``
    int64 numpy (raw labels 1..4).  This is what the on-device dataset keeps instead of the
    6x-expanded window array."""

    def __init__(self, sid, streams, starts, labels, window, channel_names):
        self.sid, self.streams, self.labels, self.window = sid, streams, labels, window
        self.starts_host = starts
        # pinned + non_blocking: a pageable copy would make the host wait for everything enqueued before it (the whole subject)
        self.starts = torch.from_numpy(starts).pin_memory().to(streams.device, non_blocking=True) if streams.is_cuda and len(starts) \
            else torch.from_numpy(starts).to(streams.device)
        self.channel_names = list(channel_names)

    def _ptr_array(self, idx):
        rows = [self.streams[i] for i in idx]
        return (C.c_void_p * len(rows))(*[r.data_ptr() for r in rows]), rows

    def windows_f64(self, channel_idx=None) -> torch.Tensor:
        """``[n_win, W, C]`` float64 on the device -- the array preprocess.py:218 saves."""
        lib = _ext.lib()
        idx = list(range(self.streams.shape[0])) if channel_idx is None else list(channel_idx)
        arr, _keep = self._ptr_array(idx)
        n_win = len(self.labels)
        out = torch.empty(n_win, self.window, len(idx), dtype=torch.float64, device=self.streams.device)
        check(lib.mms_window_gather(arr, len(idx), self.streams.shape[1], ptr(self.starts), n_win, self.window, 0,
                                    None, None, None, ptr(out), stream()))
        return out


def preprocess_subject(sid, data, protocol, target_fs=None, include_wrist=False, device=None) -> SubjectStreams:
    """One iteration of the reference subject loop (preprocess.py:138-200) up to, but not
    including, the materialisation of the windows."""
    target_fs = RAW_FS if target_fs is None else target_fs
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    chest = {k.decode('utf-8') if isinstance(k, bytes) else k: v for k, v in data[b'signal'][b'chest'].items()}
    groups = [_UPLOADER._columns(chest, CHEST_CHANNELS)]
    if include_wrist:
        wrist = {k.decode('utf-8') if isinstance(k, bytes) else k: v for k, v in data[b'signal'][b'wrist'].items()}
        groups += [_UPLOADER._columns(wrist, [name]) for name in WRIST_CHANNELS]
    staged = _UPLOADER.groups(groups, device)          # the whole subject: one staging pass, one H2D copy
    rows = staged[0]
    num = resampled_length(rows.shape[1], ORIGINAL_CHEST_FS, target_fs)
    streams = resample_subject_rows(rows, dict(zip(WRIST_CHANNELS, staged[1:])) if include_wrist else None, target_fs)
    names = list(CHEST_CHANNEL_NAMES) + (list(WRIST_CHANNEL_NAMES) if include_wrist else [])
    starts, labels, window = window_plan(protocol, target_fs)
    if len(starts) and starts.max() + window > num:
        raise ValueError(f"{sid}: a window runs past the end of the resampled stream")
    return SubjectStreams(sid, streams, starts, labels, window, names)


class _NpyWriter:
    """``np.save(path, array)`` (reference preprocess.py:217-218) for a DEVICE tensor without the two extra host copies of
    ``tensor.cpu().numpy()`` + ``np.save``: the tensor is copied once into a reusable pinned buffer and the ``.npy`` file
    (format 1.0 header + C-order payload, byte-identical to what ``np.save`` writes) is streamed from that buffer
    (SURVEY §8f N4).  ``dtype`` may down-convert on the device first (e.g. ``torch.float32`` for a half-size file; the
    reference's format is float64, which stays the default everywhere)."""

    def __init__(self):
        self._pinned = None

    def save(self, path, tensor: torch.Tensor, dtype=None):
        t = tensor.detach().contiguous()
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        nbytes = t.numel() * t.element_size()
        if t.is_cuda:
            if self._pinned is None or self._pinned.numel() < nbytes:
                self._pinned = torch.empty(int(nbytes * 1.25) + 64, dtype=torch.uint8).pin_memory()
            host = self._pinned[:nbytes].view(t.dtype).view(t.shape)
            host.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            host = t
        arr = host.numpy()
        with open(path, 'wb') as f:
            np.lib.format.write_array_header_1_0(f, {'descr': np.lib.format.dtype_to_descr(arr.dtype), 'fortran_order': False,
                                                     'shape': tuple(arr.shape)})
            if arr.size:
                f.write(memoryview(arr.reshape(-1)).cast('B'))
        return nbytes


_NPY_WRITER = _NpyWriter()


def preprocess_subjects_sharded(items, target_fs=None, include_wrist=False, group=None, device=None, subject_fn=None):
    """Preprocess a list of subjects with the work sharded over the ranks of a process group, then give every rank
    every subject's resampled streams (each LOSO fold trains on 11 subjects, validates on 3, tests on 1 -- all 15
    are needed everywhere).

    ``items``: list of ``(sid, data, protocol)``; ``data`` is the unpickled recording or a zero-argument callable
    that loads it (only the owning rank calls it).  Subject ``i`` is resampled by rank ``i % world``; the streams
    (``[n_channels, num]`` float64, ~43 MB per subject) then travel GPU-to-GPU with one NCCL broadcast per subject
    over NVLink instead of every rank pushing all raw recordings (269 MB per subject) through its own PCIe link.
    Without an initialised process group this is a plain loop.  Returns ``{sid: SubjectStreams}``.
    ``device`` / ``subject_fn`` (default: the current CUDA device / ``preprocess_subject``) exist so that the exchange
    logic can be exercised over gloo on a CPU-only box; the default path has no CPU fallback."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    target_fs = RAW_FS if target_fs is None else target_fs
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    subject_fn = preprocess_subject if subject_fn is None else subject_fn
    mine = {}
    for i, (sid, data, protocol) in enumerate(items):
        if i % world == rank:
            mine[sid] = subject_fn(sid, data() if callable(data) else data, protocol, target_fs, include_wrist=include_wrist,
                                   device=device)
    if world == 1:
        return mine
    meta = [None] * world
    dist.all_gather_object(meta, {sid: (tuple(s.streams.shape), s.starts_host, s.labels, s.window, s.channel_names)
                                  for sid, s in mine.items()}, group=group)
    out = {}
    for i, (sid, _, _) in enumerate(items):
        owner = i % world
        shape, starts, labels, window, names = meta[owner][sid]
        streams = mine[sid].streams if owner == rank else torch.empty(shape, dtype=torch.float64, device=device)
        dist.broadcast(streams, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        out[sid] = mine[sid] if owner == rank else SubjectStreams(sid, streams, starts, labels, window, names)
    return out


def run_preprocessing(wesad_root=None, output_path=None, subject_ids=None, include_wrist=False):
    """reference preprocess.py:126-242 for ``PROCESS_TARGETS = ['raw']``."""
    wesad_root = Path(WESAD_ROOT if wesad_root is None else wesad_root)
    output_path = Path(OUTPUT_PATH if output_path is None else output_path)
    for target in PROCESS_TARGETS:
        if target != 'raw':
            raise NotImplementedError(f"PROCESS_TARGETS entry {target!r}: only 'raw' is in scope "
                                      "(the 'feature' / 'raw-align' branches need neurokit2)")
    subject_ids = [f"S{i}" for i in range(2, 18) if i != 12] if subject_ids is None else subject_ids
    raw_path = output_path / ('all_raw' if include_wrist else 'chest_raw')
    raw_path.mkdir(parents=True, exist_ok=True)
    names = CHEST_CHANNEL_NAMES + (WRIST_CHANNEL_NAMES if include_wrist else [])
    with open(raw_path / '_channel_names.txt', 'w') as f:
        for name in names:
            f.write(f"{name}\n")
    # RAW_FS defaults to 64 Hz here (north star) where the reference ships 128 (preprocess.py:21): say so at run time and leave
    # a sidecar next to the windows, so that arrays of different rates in the same ./data tree can be told apart
    window = int(RAW_WINDOW_SEC * RAW_FS)
    print(f"Resampling to RAW_FS = {RAW_FS} Hz: windows of {window} samples x {len(names)} channels "
          f"(the reference's default RAW_FS = 128 gives 7680)")
    import json
    (raw_path / '_preprocess_meta.json').write_text(json.dumps({"raw_fs": RAW_FS, "window_samples": window, "stride_samples": int(RAW_STRIDE_SEC * RAW_FS),
                                                                 "channels": names, "producer": "multimodalsignal_b200.preprocess"}))
    done = []
    for sid in subject_ids:
        data = load_pkl(sid, wesad_root)
        if data is None:
            continue
        protocol = parse_quest_csv(sid, wesad_root)
        sub = preprocess_subject(sid, data, protocol, RAW_FS, include_wrist=include_wrist)
        if len(sub.labels):
            X = sub.windows_f64()                                   # [n_win, W, C] float64 on the device
            _NPY_WRITER.save(raw_path / f'{sid}_X.npy', X)          # == np.save of the same array (preprocess.py:217)
            np.save(raw_path / f'{sid}_y.npy', sub.labels)
            print(f"  - {sid} (raw): Saved {len(sub.labels)} windows. Raw shape: {tuple(X.shape)}")
            done.append(sid)
    print("\nPreprocessing complete.")
    return done


if __name__ == '__main__':
    run_preprocessing()
