"""Host-side logic of the drop-in modules (no GPU): index arithmetic, protocol parsing, fold
tables, fold sharding over ranks (gloo, world_size 2), EarlyStopping semantics, file dataset."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden
from multimodalsignal_b200 import synth
from oracle import preprocess_oracle as po


@pytest.mark.parametrize("fs", [64, 128])
def test_window_plan_bit_exact_vs_reference(fs):
    """Window start indices and labels of the drop-in == the reference's own loop (golden, full
    15-subject protocol with float-trap start times)."""
    from multimodalsignal_b200 import preprocess as pp
    z, _ = load_golden("preprocess_golden.npz")
    for sid in synth.ALL_SUBJECTS:
        protocol = po.apply_subject_quirk(sid, synth.FULL_PROTOCOL)
        starts, labels, w = pp.window_plan(protocol, fs)
        assert starts.dtype == np.int64 and labels.dtype == np.int64
        assert np.array_equal(starts, z[f"full/{fs}/{sid}/starts"]), sid
        assert np.array_equal(labels, z[f"full/{fs}/{sid}/labels"]), sid
        assert w == 60 * fs


def test_resampled_length_matches_reference_expression():
    from multimodalsignal_b200 import preprocess as pp
    for n in [1, 699, 700, 4200000, 4200137, 4201918, 7999999]:
        for f0, f1 in [(700, 64), (700, 128), (32, 64), (4, 64), (64, 64)]:
            assert pp.resampled_length(n, f0, f1) == int(n * (f1 / f0))


def test_parse_quest_csv_and_quirk(tmp_path):
    from multimodalsignal_b200 import preprocess as pp
    for sid in ("S2", "S5"):
        d = tmp_path / sid
        d.mkdir()
        (d / f"{sid}_quest.csv").write_text(synth.quest_csv_text(synth.FULL_PROTOCOL))
    p5 = pp.parse_quest_csv("S5", tmp_path)
    assert [t for t, _, _ in p5] == [t for t, _, _ in synth.FULL_PROTOCOL]
    assert p5[1][1] == 32.05 and p5[1][2] == 43.5
    p2 = pp.parse_quest_csv("S2", tmp_path)
    assert p2[0] == ("Base", (5.0 + 25.0) / 2, 25.0)            # preprocess.py:53-57
    assert p2[1:] == p5[1:]
    assert pp.load_pkl("S99", tmp_path) is None                  # preprocess.py:66-68


def test_constants_match_reference_contract():
    from multimodalsignal_b200 import main as mm, preprocess as pp
    assert pp.CHEST_CHANNEL_NAMES == ["chest_ACC_x", "chest_ACC_y", "chest_ACC_z", "chest_ECG", "chest_EDA",
                                      "chest_EMG", "chest_Resp", "chest_Temp"]
    assert pp.TASK_TO_LABEL_MAP == {'Base': 1, 'TSST': 2, 'Fun': 3, 'Medi1': 4, 'Medi2': 4}
    assert (pp.RAW_WINDOW_SEC, pp.RAW_STRIDE_SEC, pp.ORIGINAL_CHEST_FS) == (60, 10, 700)
    assert mm.ALL_SUBJECTS == synth.ALL_SUBJECTS and len(mm.ALL_SUBJECTS) == 15
    assert (mm.SEED, mm.EPOCHS, mm.BATCH_SIZE, mm.LEARNING_RATE, mm.PATIENCE, mm.WEIGHTS_DECAY) == (42, 100, 64, 0.001, 20, 1e-4)
    assert mm.CHANNELS_TO_USE == ['chest_ECG', 'chest_EDA', 'chest_Resp']
    cfg = mm.trainer_config()['trainer']
    assert cfg['early_stopping'] == {'enabled': True, 'patience': 20, 'delta': 0}


def test_fold_split_matches_reference_table():
    from multimodalsignal_b200 import main as mm
    table = json.loads((GOLDEN / "fold_table.json").read_text())["folds"]
    for s in mm.ALL_SUBJECTS:
        tr, va = mm.fold_split(s)
        assert tr == table[s]["train"] and va == table[s]["val"], s


def test_folds_for_rank_partition():
    from multimodalsignal_b200.main import folds_for_rank
    for world in (1, 2, 4, 8, 15, 16):
        seen = sorted(f for r in range(world) for f in folds_for_rank(15, r, world))
        assert seen == list(range(15))
        assert max(len(folds_for_rank(15, r, world)) for r in range(world)) == -(-15 // world)


def test_early_stopping_is_bug_compatible(tmp_path):
    """with val_loss as the score the reference saves when the loss does NOT improve."""
    from multimodalsignal_b200.trainer import EarlyStopping
    model = torch.nn.Linear(2, 2)
    saves = []
    es = EarlyStopping(patience=2, delta=0, checkpoint_path=tmp_path / "best_model.pt")
    es.save_checkpoint = lambda m: saves.append(es.best_score)
    es(1.0, model)            # first call always saves
    es(0.9, model)            # loss improved -> reference counts patience (1/2), no save
    assert es.counter == 1 and len(saves) == 1
    es(1.2, model)            # loss got worse -> reference saves and resets
    assert es.counter == 0 and len(saves) == 2 and es.best_score == 1.2
    es(1.1, model); es(1.0, model)
    assert es.early_stop
    fixed = EarlyStopping(patience=2, delta=0, checkpoint_path=tmp_path / "b.pt", fixed=True)
    fixed.save_checkpoint = lambda m: saves.append("fixed")
    fixed(1.0, model); fixed(0.9, model)
    assert fixed.counter == 0 and saves[-1] == "fixed"      # the intended behaviour: improvement saves


def test_file_dataset_matches_reference(tmp_path):
    """``WesadDataset`` (file path, host numpy) vs reference dataset.py output (golden)."""
    from multimodalsignal_b200.dataset import WesadDataset
    z, _ = load_golden("preprocess_golden.npz")
    names = bytes(z["channel_names"]).decode().split()
    for sid in ("S2", "S5"):
        idx = synth.ALL_SUBJECTS.index(sid)
        sub = synth.make_subject(sid, idx, minutes=synth.SHORT_MINUTES, protocol=synth.SHORT_PROTOCOL, with_wrist=False)
        X, y = po.preprocess_subject(sid, sub.chest, sub.protocol, 64)
        np.save(tmp_path / f"{sid}_X.npy", X)
        np.save(tmp_path / f"{sid}_y.npy", y)
    chans = ["chest_ECG", "chest_EDA", "chest_EMG", "chest_Resp"]
    for mode in ("stress_binary", "ternary"):
        ds = WesadDataset(tmp_path, ["S2", "S5", "S99"], chans, names, classification_mode=mode)
        assert len(ds) == int(z[f"dataset/{mode}/len"])
        assert np.array_equal(ds.labels, z[f"dataset/{mode}/labels"])
        np.testing.assert_allclose(ds.data[:, ::61, :], z[f"dataset/{mode}/data_sub"], atol=1e-9)
        x3, y3 = ds[3]
        assert x3.dtype == torch.float32 and tuple(x3.shape) == (4, 3840) and y3.dtype == torch.int64
        np.testing.assert_allclose(x3.numpy()[:, ::61], z[f"dataset/{mode}/item3_x_sub"], atol=1e-6)
    with pytest.raises(ValueError):
        WesadDataset(tmp_path, ["S2"], chans, names, classification_mode="amusement_binary")
    with pytest.raises(ValueError):
        WesadDataset(tmp_path, ["S98", "S99"], chans, names)                                     # dataset.py:53-54


def test_fold_sharding_world_size_2_gloo(tmp_path):
    """N > 1 path on CPU: two gloo ranks shard the 15 folds, rank 0 writes cv_summary.txt."""
    script = tmp_path / "shard.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {str(ROOT)!r})
        import torch.distributed as dist
        from multimodalsignal_b200 import main as mm
        dist.init_process_group("gloo")
        def fake_fold(fold_index, subject, out_dir, names, streams):
            return {{'subject': subject, 'accuracy': 0.5 + 0.01 * fold_index, 'f1_score': 0.4 + 0.01 * fold_index,
                    'rank': dist.get_rank()}}
        res = mm.run_simple_experiment({str(tmp_path)!r}, None, [], fold_fn=fake_fold)
        assert [r['subject'] for r in res] == mm.ALL_SUBJECTS
        assert sorted(set(r['rank'] for r in res)) == [0, 1]
        assert [r['rank'] for r in res] == [i % 2 for i in range(15)]
        dist.destroy_process_group()
    """))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    text = (tmp_path / "cv_summary.txt").read_text(encoding="utf-8")
    assert "测试 S2: Accuracy = 0.5000, F1-score = 0.4000" in text
    assert "测试 S17: Accuracy = 0.6400" in text
    assert "平均准确率 (Accuracy): 0.5700" in text


@pytest.mark.parametrize("nc", [2, 3, 5])
def test_metrics_from_confusion_equal_sklearn(nc):
    """accuracy / weighted F1 from the on-device confusion matrix == sklearn's accuracy_score and
    f1_score(average='weighted') on the prediction lists (reference trainer.py:229-230), including classes
    that never occur or are never predicted."""
    import warnings
    from sklearn.metrics import accuracy_score, f1_score
    from multimodalsignal_b200.trainer import metrics_from_confusion
    rng = np.random.default_rng(nc)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        y = rng.integers(0, nc, n)
        p = rng.integers(0, nc, n)
        if trial % 4 == 1:
            p[:] = 0                              # one class predicted only
        if trial % 4 == 2:
            y[y == nc - 1] = 0                    # a class without support
        if trial % 4 == 3:
            p = y.copy()                          # perfect
        conf = np.zeros((nc, nc), dtype=np.int64)
        np.add.at(conf, (y, p), 1)
        acc, f1 = metrics_from_confusion(conf)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert abs(acc - accuracy_score(y, p)) <= 1e-15
            assert abs(f1 - f1_score(y, p, average='weighted')) <= 1e-12
    assert metrics_from_confusion(np.zeros((nc, nc))) == (0.0, 0.0)


def test_run_folds_interleaved_schedules_generators_by_event_completion():
    """The single-thread fold scheduler (main.run_folds_interleaved): at most K generators are live, a generator is resumed
    only once the event it yielded reports completion, a finished fold frees its stream for the next pending one, and the
    results come back in fold order.  CUDA events / streams are replaced by stand-ins (no GPU here)."""
    from unittest import mock
    from multimodalsignal_b200 import main as mm

    class Ev:
        def __init__(self, polls):
            self.polls = polls

        def query(self):
            self.polls -= 1
            return self.polls <= 0

    live, max_live, order = set(), [0], []

    def fold(i, stream):
        live.add(i)
        max_live[0] = max(max_live[0], len(live))
        for step in range(3 + i % 2):
            ev = Ev(polls=1 + (i * 7 + step) % 4)
            yield ev
            assert ev.polls <= 0, "resumed before its event completed"
        live.discard(i)
        order.append(i)
        return {"fold": i, "stream": stream}

    with mock.patch.object(mm.torch.cuda, "Stream", side_effect=lambda: object()):
        res = mm.run_folds_interleaved([0, 1, 2, 3, 4, 5, 6], fold, concurrent_folds=3)
    assert [r["fold"] for r in res] == [0, 1, 2, 3, 4, 5, 6]
    assert max_live[0] == 3 and sorted(order) == list(range(7))
    assert all("start_s" in r and r["end_s"] >= r["start_s"] for r in res)
    assert len({id(r["stream"]) for r in res}) == 3          # three streams shared by seven folds


@pytest.mark.parametrize("shape,dtype", [((7, 5, 3), torch.float64), ((0, 4, 8), torch.float64), ((3, 11), torch.float32), ((5,), torch.int64)])
def test_npy_writer_is_byte_identical_to_np_save(tmp_path, shape, dtype):
    """The direct .npy writer of the preprocess path (header + payload streamed from one staging buffer) produces the
    very bytes ``np.save`` writes for the same array (reference preprocess.py:217-218), and ``np.load`` reads it back."""
    from multimodalsignal_b200.preprocess import _NpyWriter
    g = torch.Generator().manual_seed(sum(shape) + 1)
    t = (torch.randn(shape, generator=g, dtype=torch.float64) * 100).to(dtype)
    _NpyWriter().save(tmp_path / "a.npy", t)
    np.save(tmp_path / "b.npy", t.numpy())
    assert (tmp_path / "a.npy").read_bytes() == (tmp_path / "b.npy").read_bytes()
    back = np.load(tmp_path / "a.npy")
    assert back.dtype == t.numpy().dtype and back.shape == tuple(shape) and np.array_equal(back, t.numpy())
    if dtype == torch.float64:            # the optional half-size on-disk variant
        _NpyWriter().save(tmp_path / "c.npy", t, dtype=torch.float32)
        assert np.array_equal(np.load(tmp_path / "c.npy"), t.numpy().astype(np.float32))


def test_sharded_preprocess_exchange_world_size_2_gloo(tmp_path):
    """N > 1 path of ``preprocess_subjects_sharded`` on CPU (gloo, two ranks): subject i is produced by rank i % 2 only
    (the loader callable of the other rank's subjects is never invoked), every rank ends up with every subject's streams,
    window plan and labels, bit for bit.  The per-subject GPU work is replaced by a stand-in (no GPU here)."""
    script = tmp_path / "shard_pre.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {str(ROOT)!r})
        import numpy as np, torch
        import torch.distributed as dist
        from multimodalsignal_b200 import preprocess as pp
        dist.init_process_group("gloo")
        rank = dist.get_rank()
        loaded = []
        def loader(i):
            def f():
                loaded.append(i)
                return {{"i": i}}
            return f
        def fake_subject(sid, data, protocol, target_fs, include_wrist=False, device=None):
            i = data["i"]
            streams = (torch.arange(3 * (50 + i), dtype=torch.float64).view(3, 50 + i) + 1000 * i).to(device)
            starts = np.arange(0, 10 + i, 5, dtype=np.int64)
            labels = (np.arange(len(starts)) % 4 + 1).astype(np.int64)
            return pp.SubjectStreams(sid, streams, starts, labels, 20, ["a", "b", "c"])
        items = [(f"S{{i}}", loader(i), []) for i in range(5)]
        out = pp.preprocess_subjects_sharded(items, 64, device=torch.device("cpu"), subject_fn=fake_subject)
        assert loaded == [i for i in range(5) if i % 2 == rank], loaded
        assert list(out) == [f"S{{i}}" for i in range(5)]
        for i in range(5):
            s = out[f"S{{i}}"]
            assert torch.equal(s.streams, torch.arange(3 * (50 + i), dtype=torch.float64).view(3, 50 + i) + 1000 * i)
            assert np.array_equal(s.starts_host, np.arange(0, 10 + i, 5)) and s.window == 20 and s.channel_names == ["a", "b", "c"]
            assert np.array_equal(s.labels, np.arange(len(s.starts_host)) % 4 + 1)
        dist.barrier()
        if rank == 0:
            print("sharded ok")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded ok" in r.stdout


def test_window_plan_edge_cases_equal_oracle():
    """Host index arithmetic on irregular protocols -- unknown task names (skipped, reference preprocess.py:163-164), task
    names with inner spaces ('Medi 1'), segments shorter than one window (no window), segments of exactly one window,
    two-decimal start times that hit the float64 truncation trap -- is bit-identical to the oracle (itself pinned to the
    reference's loop by tests/test_oracle_preprocess.py)."""
    from multimodalsignal_b200 import preprocess as pp
    rng = np.random.default_rng(7)
    names = ["Base", "TSST", "Medi 1", "Fun", "Medi 2", "sRead", "fRead", "bRead", " Base ", "Medi1"]
    for trial in range(200):
        protocol = []
        t = 0.0
        for _ in range(int(rng.integers(0, 7))):
            start = round(t + float(rng.uniform(0.0, 3.0)), 2)
            dur = [0.2, 0.99, 1.0, 1.01, float(rng.uniform(0.0, 12.0))][int(rng.integers(0, 5))]
            end = round(start + dur, 2)
            protocol.append((names[int(rng.integers(0, len(names)))], start, end))
            t = end
        for fs in (64, 128):
            s1, l1, w1 = pp.window_plan(protocol, fs)
            s2, l2, w2 = po.window_plan(protocol, fs)
            assert w1 == w2 == 60 * fs
            assert s1.dtype == np.int64 and l1.dtype == np.int64
            assert np.array_equal(s1, s2) and np.array_equal(l1, l2), (protocol, fs)
    assert pp.window_plan([], 64)[0].shape == (0,)
    assert pp.window_plan([("sRead", 1.0, 30.0)], 64)[0].shape == (0,)            # unknown task: no windows
    assert pp.window_plan([("Base", 5.0, 5.99)], 64)[0].shape == (0,)             # shorter than 60 s: no window
    assert len(pp.window_plan([("Base", 5.0, 6.0)], 64)[0]) == 1                  # exactly one window


def test_map_labels_modes_and_error():
    """reference dataset.py:28-34."""
    from multimodalsignal_b200.dataset import map_labels
    raw = np.array([1, 2, 3, 4, 2, 1])
    assert map_labels(raw, "stress_binary").tolist() == [0, 1, 0, 0, 1, 0]
    assert map_labels(raw, "ternary").tolist() == [0, 2, 1, 0, 2, 0]
    assert map_labels(raw, "stress_binary").tolist() == po.map_labels(raw, "stress_binary").tolist()
    assert map_labels(raw, "ternary").tolist() == po.map_labels(raw, "ternary").tolist()
    with pytest.raises(ValueError):
        map_labels(raw, "quaternary")
