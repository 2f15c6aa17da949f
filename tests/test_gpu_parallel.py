"""Intra-fold data parallelism (SyncBN + one flat gradient reduction), emulated on ONE GPU: two ranks'
phase-split steps run in lockstep with their sync tensors summed by hand -- exactly what the NCCL
all-reduce does -- and must reproduce the single-device full-batch step (the reference computes BN
statistics over the whole batch, models.py:47,51)."""
import copy

import numpy as np
import pytest
import torch

from conftest import golden_state, load_golden

pytestmark = pytest.mark.gpu


def _model(meta, sd):
    from multimodalsignal_b200.models import CnnGruAttentionModel
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = CnnGruAttentionModel(meta["C"], meta["num_classes"], dropout=0.0, **meta["kwargs"])
    m.load_state_dict(sd, strict=True)
    return m.cuda().train()


@pytest.mark.parametrize("case,splits", [("c6_t640", (2, 2)), ("c6_t640", (1, 3)), ("c8_h32_l1", (2, 1))])
def test_two_rank_data_parallel_equals_full_batch(case, splits):
    from multimodalsignal_b200.parallel import DataParallelTrainStep, emulate_ranks
    from multimodalsignal_b200.trainer import FlatAdam, FusedTrainStep
    z, meta = load_golden(f"model_{case}.npz")
    sd = golden_state(z, "sd")
    x, y = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["y"]).cuda()
    B, T = x.shape[0], x.shape[2]
    assert sum(splits) == B
    # single device, full batch (already pinned to the reference's Adam trajectory by test_gpu_model.py)
    ref = _model(meta, sd)
    ref_step = FusedTrainStep(ref, FlatAdam(ref, lr=1e-3, weight_decay=1e-4), B, T, use_graph=False)
    # two emulated ranks
    ranks, steps, lo = [], [], 0
    for r, b in enumerate(splits):
        m = _model(meta, sd)
        ranks.append(m)
        steps.append(DataParallelTrainStep(m, FlatAdam(m, lr=1e-3, weight_decay=1e-4), b, B, T, rank=r))
    losses = []
    for it in range(3):
        ref_step(x, y)
        lo, batches = 0, []
        for b in splits:
            batches.append((x[lo:lo + b], y[lo:lo + b]))
            lo += b
        emulate_ranks(steps, batches)
        losses.append(sum(float(s.loss.item()) for s in steps))
        assert abs(losses[-1] - ref_step.last_loss()) < 2e-5
    ref_sd = ref.state_dict()
    for m in ranks:
        for k, v in m.state_dict().items():
            if v.numel() == 0:
                continue
            np.testing.assert_allclose(v.float().cpu().numpy(), ref_sd[k].float().cpu().numpy(), atol=2e-5, err_msg=k)
    if "adam_losses" in z.files:        # and therefore also the reference's own trajectory
        np.testing.assert_allclose(losses[:int(z["adam_steps"])], z["adam_losses"][:3], atol=1e-4)


def test_nccl_data_parallel_two_gpus():
    """The same check over real NCCL (skipped on a single-GPU box)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633", str(ROOT / "tools" / "dp_check.py")],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dp_check ok" in r.stdout


def test_sharded_preprocess_two_gpus():
    """Subject-sharded resampling + NCCL broadcast of the streams == single-process preprocessing, bit for bit
    (skipped on a single-GPU box)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29637", str(ROOT / "tools" / "shard_check.py")],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "shard_check ok" in r.stdout
