"""Device-resident epoch plumbing on a B200: ``mms_batch_gather`` (shuffled batch assembly through a device
cursor), ``mms_eval_accumulate`` (loss / arg-max / confusion counts) and the ``Trainer`` paths built on them,
checked against torch on the same inputs and against the host-fed ``Trainer`` path (which
tests/test_gpu_trainer.py pins to the reference)."""
import ctypes as C
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from multimodalsignal_b200 import _ext
    return _ext.lib()


def P(t):
    return None if t is None else t.data_ptr()


def ST():
    return torch.cuda.current_stream().cuda_stream


def ok(rc):
    from multimodalsignal_b200 import _ext
    _ext.check(rc)


@pytest.mark.parametrize("N,row,B", [(37, 8, 5), (300, 6 * 640, 64), (129, 3 * 3840, 64), (64, 4, 64)])
def test_batch_gather_walks_the_permutation(lib, N, row, B):
    """Bit-exact copy of the selected rows; the device cursor advances by B per launch; a ragged tail batch
    is a launch with a smaller B (DataLoader drop_last=False)."""
    g = torch.Generator().manual_seed(N)
    data = torch.randn(N, row, generator=g).cuda()
    labels = torch.randint(0, 4, (N,), generator=g).cuda()
    perm_h = torch.randperm(N, generator=g)
    perm = torch.cat([perm_h, torch.zeros(B, dtype=torch.int64)]).cuda()
    cursor = torch.zeros(1, dtype=torch.int64, device="cuda")
    scratch = torch.zeros(1, dtype=torch.int32, device="cuda")
    full, tail = divmod(N, B)
    pos = 0
    for b in [B] * full + ([tail] if tail else []):
        x = torch.full((b, row), float("nan"), device="cuda")
        y = torch.full((b,), -1, dtype=torch.int64, device="cuda")
        ok(lib.mms_batch_gather(P(data), P(labels), P(perm), P(cursor), N, row, b, P(x), P(y), 1, P(scratch), ST()))
        sel = perm_h[pos:pos + b].cuda()
        assert torch.equal(x, data[sel]) and torch.equal(y, labels[sel])
        pos += b
        assert int(cursor.item()) == pos and int(scratch.item()) == 0
    # identity permutation / no cursor / no advance
    b = min(B, N)
    x = torch.empty(b, row, device="cuda")
    ok(lib.mms_batch_gather(P(data), None, None, None, N, row, b, P(x), None, 0, None, ST()))
    assert torch.equal(x, data[:b])


def test_batch_gather_rejects_bad_arguments(lib):
    data = torch.zeros(4, 6, device="cuda")
    x = torch.zeros(2, 6, device="cuda")
    assert lib.mms_batch_gather(P(data), None, None, None, 4, 6, 2, P(x), None, 0, None, ST()) < 0       # row not a multiple of 4 floats
    data = torch.zeros(4, 8, device="cuda")
    x = torch.zeros(2, 8, device="cuda")
    assert lib.mms_batch_gather(P(data), None, None, None, 4, 8, 2, P(x), None, 1, None, ST()) < 0       # advance without a cursor


@pytest.mark.parametrize("B,nc", [(1, 2), (64, 2), (300, 3), (1000, 8)])
def test_eval_accumulate_vs_torch(lib, B, nc):
    g = torch.Generator().manual_seed(B + nc)
    logits = (torch.randn(B, nc, generator=g) * 3).cuda()
    if B > 4:
        logits[3] = logits[3, 0]                  # an exact tie: first maximum wins (torch.argmax)
    labels = torch.randint(0, nc, (B,), generator=g).cuda()
    preds = torch.full((B,), -1, dtype=torch.int64, device="cuda")
    conf = torch.zeros(nc * nc, dtype=torch.int64, device="cuda")
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    for rep in range(2):                          # accumulates across calls
        ok(lib.mms_eval_accumulate(P(logits), P(labels), B, nc, P(preds), P(conf), P(loss), ST()))
    ref_pred = torch.argmax(torch.softmax(logits, dim=1), dim=1) if B <= 4 else torch.argmax(logits, dim=1)
    assert torch.equal(preds, ref_pred)
    ref_conf = torch.zeros(nc, nc, dtype=torch.int64)
    for yy, pp in zip(labels.cpu().tolist(), ref_pred.cpu().tolist()):
        ref_conf[yy, pp] += 2
    assert torch.equal(conf.cpu().view(nc, nc), ref_conf)
    ref_loss = 2 * torch.nn.functional.cross_entropy(logits.double(), labels, reduction="sum").item()
    assert abs(loss.item() - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))      # fp32 per-row terms, float64 sum


EPOCH_RE = re.compile(r"训练损失: ([\d.]+) \| 验证损失: ([\d.]+) \| 验证Acc: ([\d.]+) \| 验证F1: ([\d.]+)")


def _run_trainer(tmp, tag, loaders, seed, epochs=3):
    import warnings
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import Trainer
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = CnnGruAttentionModel(in_channels=3, num_classes=2, dropout=0.0)
    cfg = {'trainer': {'epochs': epochs, 'learning_rate': 1e-3,
                       'early_stopping': {'enabled': True, 'patience': 20, 'delta': 0}, 'weight_decay': 1e-4}}
    tr = Trainer(model, tmp / tag, cfg)
    tl, vl, te = loaders
    tr.train(tl, vl)
    test = tr.evaluate(te, is_test=True)
    log = (tmp / tag / "training_log.txt").read_text(encoding="utf-8")
    return model, tr, test, [tuple(float(v) for v in m.groups()) for m in EPOCH_RE.finditer(log)]


class _TensorSet(torch.utils.data.Dataset):
    """Host-side view of a DeviceWesadDataset-like object: what the reference-style DataLoader path consumes."""

    def __init__(self, data, labels):
        self.data, self.labels = data.cpu(), labels.cpu()

    def __len__(self):
        return len(self.labels)

    def __getitem__(self, i):
        return self.data[i], self.labels[i]


class _DevSet:
    def __init__(self, data, labels):
        self.data, self.labels = data.cuda(), labels.cuda()

    def __len__(self):
        return int(self.labels.shape[0])


def test_device_epoch_equals_host_fed_epochs(tmp_path):
    """Same data, same order (shuffle off), dropout 0: the device-cursor path (gather inside the graph, ragged tail
    batch, chunked graph evaluation, confusion-matrix metrics, asynchronous best_model.pt) reproduces the host-fed
    DataLoader path up to the run-to-run noise of the fp32 atomics in the weight-gradient reductions (amplified by Adam on
    near-zero gradients): losses to 5e-4, accuracy / F1 up to two borderline windows, parameters and best_model.pt to 1e-3."""
    from torch.utils.data import DataLoader
    from multimodalsignal_b200.dataset import DeviceBatchLoader
    g = torch.Generator().manual_seed(5)
    n_tr, n_va, n_te, Cc, T, B = 150, 70, 33, 3, 640, 64       # 150 = 2 full batches + a tail of 22
    def mk(n):
        y = torch.randint(0, 2, (n,), generator=g)
        x = torch.randn(n, Cc, T, generator=g) + y[:, None, None].float() * torch.sin(torch.arange(T) / 7.0)
        return x, y
    sets = [mk(n) for n in (n_tr, n_va, n_te)]
    host = [DataLoader(_TensorSet(x, y), batch_size=B, shuffle=False) for x, y in sets]
    devl = [DeviceBatchLoader(_DevSet(x, y), B, shuffle=False) for x, y in sets]
    m_h, tr_h, test_h, ep_h = _run_trainer(tmp_path, "host", host, seed=11)
    m_d, tr_d, test_d, ep_d = _run_trainer(tmp_path, "dev", devl, seed=11)
    assert len(ep_h) == len(ep_d) == 3
    for (a, b, c, d), (e, f, gg, h) in zip(ep_h, ep_d):
        # two GPU runs differ by the ordering of fp32 atomics, which Adam amplifies on near-zero gradients: losses to 5e-4 (the
        # log prints 4 decimals), accuracy / F1 up to two borderline windows of the 70-window validation set
        assert abs(a - e) <= 5e-4 and abs(b - f) <= 5e-4 and abs(c - gg) <= 2.0 / n_va + 1e-4 and abs(d - h) <= 0.06
    assert abs(test_h[0] - test_d[0]) <= 5e-4 and abs(test_h[1] - test_d[1]) <= 2.0 / n_te + 1e-9 and abs(test_h[2] - test_d[2]) <= 0.1
    assert tr_h.windows_trained == tr_d.windows_trained == 3 * n_tr
    torch.testing.assert_close(m_h.flat_parameters(), m_d.flat_parameters(), rtol=0, atol=1e-3)   # Adam amplifies the fp32-atomic ordering noise of near-zero gradients (bound: lr * steps = 9e-3)
    ck_h = torch.load(tmp_path / "host" / "best_model.pt", weights_only=True)
    ck_d = torch.load(tmp_path / "dev" / "best_model.pt", weights_only=True)
    assert list(ck_h) == list(ck_d) and len(ck_d) == 34
    for k in ck_h:
        assert ck_h[k].shape == ck_d[k].shape and ck_h[k].dtype == ck_d[k].dtype
        torch.testing.assert_close(ck_h[k].cpu().float(), ck_d[k].cpu().float(), rtol=0, atol=1e-3)


def test_device_epoch_shuffles_like_the_loader(tmp_path):
    """With shuffle on, the device source draws ``torch.randperm`` from the CPU generator exactly as
    ``DeviceBatchLoader.__iter__`` does, so both see the same batches for the same seed."""
    from multimodalsignal_b200.dataset import DeviceBatchLoader
    from multimodalsignal_b200.trainer import DeviceBatchSource
    x = torch.arange(20 * 8, dtype=torch.float32).view(20, 2, 4)
    ds = _DevSet(x, torch.arange(20))
    loader = DeviceBatchLoader(ds, 8, shuffle=True)
    torch.manual_seed(3)
    seen = torch.cat([yb for _, yb in loader]).cpu()
    torch.manual_seed(3)
    src = DeviceBatchSource(ds, 8)
    src.start_epoch(True)
    torch.cuda.synchronize()
    assert torch.equal(src.perm[:20].cpu(), seen) and int(src.cursor.item()) == 0


def test_sync_checkpoint_option_matches_async(tmp_path):
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import CheckpointWriter
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = CnnGruAttentionModel(3, 2).cuda()
    model.flat_parameters()
    w = CheckpointWriter()
    w.submit(model, tmp_path / "a.pt")
    with torch.no_grad():
        model.classifier[3].bias.add_(1.0)                     # later changes must not leak into the snapshot
    w.flush()
    a = torch.load(tmp_path / "a.pt", weights_only=True)
    sd = model.state_dict()
    assert list(a) == list(sd)
    torch.testing.assert_close(a["classifier.3.bias"] + 1.0, sd["classifier.3.bias"].cpu())
    for k in a:
        if k != "classifier.3.bias":
            assert torch.equal(a[k], sd[k].cpu()), k
            assert a[k].dtype == sd[k].dtype and a[k].shape == sd[k].shape


def test_interleaved_folds_equal_sequential_folds(tmp_path):
    """``run_folds_interleaved`` (several trainers resumed from one host thread, each on its own CUDA stream) gives
    every fold what it gets when the folds run one after another: same epochs and -- up to the fp32-atomic ordering noise
    of the weight-gradient reductions, which Adam amplifies -- the same losses, metrics and parameters."""
    import warnings
    from multimodalsignal_b200 import main as mm
    from multimodalsignal_b200.dataset import DeviceBatchLoader
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import Trainer, drive
    g = torch.Generator().manual_seed(9)
    Cc, T, B = 3, 640, 32
    def mk(n):
        y = torch.randint(0, 2, (n,), generator=g)
        x = torch.randn(n, Cc, T, generator=g) + y[:, None, None].float() * torch.sin(torch.arange(T) / 5.0)
        return _DevSet(x, y)
    data = {i: (mk(100 + 7 * i), mk(40), mk(30)) for i in range(3)}
    cfg = {'trainer': {'epochs': 3, 'learning_rate': 1e-3, 'quiet': True,
                       'early_stopping': {'enabled': True, 'patience': 20, 'delta': 0}, 'weight_decay': 1e-4}}

    def fold(i, stream, tag):
        torch.manual_seed(100 + i)
        gen = torch.Generator().manual_seed(100 + i)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = CnnGruAttentionModel(Cc, 2, dropout=0.0)
        tr = Trainer(model, tmp_path / f"{tag}{i}", cfg, stream=stream)
        tl = DeviceBatchLoader(data[i][0], B, shuffle=True, generator=gen)
        yield from tr.train_async(tl, DeviceBatchLoader(data[i][1], B))
        out = yield from tr.evaluate_async(DeviceBatchLoader(data[i][2], B))
        log = (tmp_path / f"{tag}{i}" / "training_log.txt").read_text(encoding="utf-8")
        return out, [tuple(float(v) for v in m.groups()) for m in EPOCH_RE.finditer(log)], model.flat_parameters().clone()

    seq = [drive(fold(i, None, "seq")) for i in range(3)]
    par = mm.run_folds_interleaved([0, 1, 2], lambda i, st: fold(i, st, "par"), concurrent_folds=3)
    torch.cuda.synchronize()
    for (o1, e1, p1), (o2, e2, p2) in zip(seq, par):
        assert len(e1) == len(e2) == 3
        for a, b in zip(e1, e2):
            assert abs(a[0] - b[0]) <= 5e-4 and abs(a[1] - b[1]) <= 5e-4 and abs(a[2] - b[2]) <= 2.0 / 40 + 1e-4 and abs(a[3] - b[3]) <= 0.1
        assert abs(o1[0] - o2[0]) <= 5e-4 and abs(o1[1] - o2[1]) <= 2.0 / 30 + 1e-9 and abs(o1[2] - o2[2]) <= 0.1
        torch.testing.assert_close(p1, p2, rtol=0, atol=1e-3)       # atomics-order noise through 12 Adam steps (bound 1.2e-2)
