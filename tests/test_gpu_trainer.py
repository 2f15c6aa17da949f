"""End-to-end drop-in check on a B200: synthetic recordings -> ``preprocess.run_preprocessing`` ->
``WesadDataset`` -> ``DataLoader`` -> ``Trainer.train`` / ``evaluate`` (fused, graph-replayed
training steps) against two epochs of the UNMODIFIED reference pipeline run on CPU
(tests/golden/trainer_golden.json, made by oracle/make_golden.py with dropout = 0)."""
import json
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from multimodalsignal_b200 import synth

pytestmark = pytest.mark.gpu

EPOCH_RE = re.compile(r"Epoch (\d+)/(\d+) \| 耗时: [\d.]+s \| 训练损失: ([\d.]+) \| 验证损失: ([\d.]+) \| 验证Acc: ([\d.]+) \| 验证F1: ([\d.]+)")
TEST_RE = re.compile(r"测试损失: ([\d.]+) \| 测试Acc: ([\d.]+) \| 测试F1: ([\d.]+)")


def _epochs(log):
    return [tuple(float(v) for v in m.groups()[2:]) for m in EPOCH_RE.finditer(log)]


@pytest.fixture(scope="module")
def pipeline(tmp_path_factory):
    from torch.utils.data import DataLoader
    from multimodalsignal_b200 import preprocess as pp
    from multimodalsignal_b200.dataset import WesadDataset
    from multimodalsignal_b200.models import CnnGruAttentionModel
    from multimodalsignal_b200.trainer import Trainer
    gold = json.loads((GOLDEN / "trainer_golden.json").read_text(encoding="utf-8"))
    case = gold["case"]
    tmp = tmp_path_factory.mktemp("pipe")
    synth.write_wesad_tree(tmp / "WESAD", subjects=["S2", "S3", "S4", "S5"], minutes=synth.SHORT_MINUTES,
                           protocol=synth.SHORT_PROTOCOL, with_wrist=False)
    pp.run_preprocessing(wesad_root=tmp / "WESAD", output_path=tmp / "data")
    path = tmp / "data" / "chest_raw"
    names = (path / "_channel_names.txt").read_text().split()
    mk = lambda subs: WesadDataset(path, subs, case["channels"], names, classification_mode="stress_binary")
    torch.manual_seed(case["seed"])
    np.random.seed(case["seed"])
    train_ds, val_ds, test_ds = mk(case["train"]), mk(case["val"]), mk(case["test"])
    tl = DataLoader(train_ds, batch_size=case["batch_size"], shuffle=True, num_workers=0)
    vl = DataLoader(val_ds, batch_size=case["batch_size"], shuffle=False, num_workers=0)
    te = DataLoader(test_ds, batch_size=case["batch_size"], shuffle=False, num_workers=0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = CnnGruAttentionModel(in_channels=3, num_classes=2, dropout=0.0)
    cfg = {'trainer': {'epochs': case["epochs"], 'learning_rate': 1e-3,
                       'early_stopping': {'enabled': True, 'patience': 20, 'delta': 0}, 'weight_decay': 1e-4}}
    trainer = Trainer(model, tmp / "fold", cfg)
    trainer.train(tl, vl)
    test = trainer.evaluate(te, is_test=True)
    return gold, tmp, model, trainer, test, (len(train_ds), len(val_ds), len(test_ds))


def test_dataset_sizes_and_log_format(pipeline):
    gold, tmp, model, trainer, test, sizes = pipeline
    assert sizes == (gold["n_train"], gold["n_val"], gold["n_test"])
    log = (tmp / "fold" / "training_log.txt").read_text(encoding="utf-8")
    mine, ref = _epochs(log), _epochs(gold["log"])
    assert len(mine) == len(ref) == gold["case"]["epochs"]
    assert "--- 训练完成 --- 总训练时长:" in log and "--- 最终测试结果 (模型原始输出) ---" in log
    assert ("EarlyStopping counter: 1/20" in log) == ("EarlyStopping counter: 1/20" in gold["log"])


def test_training_trajectory_matches_reference(pipeline):
    """Same init (same RNG order), same shuffling, dropout 0: the per-epoch losses of the CUDA path
    track the reference CPU run.  Tolerance 2e-3 on losses (fp32, two epochs of Adam updates;
    the log prints 4 decimals), exact on accuracy / F1 of these tiny sets."""
    gold, tmp, model, trainer, test, _ = pipeline
    log = (tmp / "fold" / "training_log.txt").read_text(encoding="utf-8")
    for (tr, vl, va, vf), (rtr, rvl, rva, rvf) in zip(_epochs(log), _epochs(gold["log"])):
        assert abs(tr - rtr) <= 2e-3 and abs(vl - rvl) <= 2e-3
        assert abs(va - rva) <= 1e-4 and abs(vf - rvf) <= 1e-4
    loss, acc, f1 = test
    assert abs(loss - gold["test"]["loss"]) <= 2e-3
    assert abs(acc - gold["test"]["acc"]) <= 1e-6 and abs(f1 - gold["test"]["f1"]) <= 1e-6
    np.testing.assert_allclose(model.classifier[3].bias.detach().cpu().numpy(), gold["final_fc3_bias"], atol=5e-4)
    np.testing.assert_allclose(model.cnn_encoder[1].running_mean.cpu().numpy(), gold["final_bn1_running_mean"], atol=1e-4)


def test_best_model_pt_layout_and_cross_loading(pipeline):
    """best_model.pt has the reference's 34 keys / shapes and loads with strict=True
    (trainer.py:38-39,187); weights_only load works as the reference does it."""
    from multimodalsignal_b200.models import CnnGruAttentionModel
    gold, tmp, model, trainer, test, _ = pipeline
    ckpt = torch.load(tmp / "fold" / "best_model.pt", weights_only=True)
    assert list(ckpt.keys()) == gold["best_model_keys"] and len(ckpt) == 34
    for k, v in ckpt.items():
        assert list(v.shape) == gold["best_model_shapes"][k], k
        assert v.untyped_storage().nbytes() == v.numel() * v.element_size()      # plain tensors, not flat-buffer views
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        other = CnnGruAttentionModel(3, 2, dropout=0.0)
    assert not other.load_state_dict(ckpt, strict=True).missing_keys
