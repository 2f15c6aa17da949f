"""CPU port of the reference training step that runs at the reference's own speed.

TEST / BENCH INFRASTRUCTURE ONLY (``bench.py --impl reference`` and the ``cpu_baseline`` leg).

``/root/reference`` is Python and does not travel to the GPU box, so the CPU baseline there is
this port.  Unlike ``model_oracle.py`` (explicit formulas, slow Python GRU loop, used as the
checker) it dispatches the SAME ATen library kernels the reference's ``torch.nn`` modules
dispatch on CPU -- ``conv1d``, ``batch_norm``, ``max_pool1d``, the fused ``gru`` op, ``linear``,
``cross_entropy`` and ``torch.optim.Adam`` (reference models.py:45-71, trainer.py:68-69,144-149)
-- through their functional entry points, with dropout enabled exactly where the reference has
it.  ``tests/test_oracle_model.py::test_cpu_port_*`` pins it to the reference fixtures.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def _gru_flat_weights(p, layers):
    flat = []
    for layer in range(layers):
        for suffix in ("", "_reverse"):
            for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                flat.append(p[f"gru.{name}_l{layer}{suffix}"])
    return flat


def forward(p, bufs, x, *, training, dropout, layers=2, attention=True):
    """reference models.py:73-81 through functional library calls.  ``bufs`` holds the BatchNorm
    running statistics (updated in place when ``training``)."""
    if attention:
        s = F.adaptive_avg_pool1d(x, 1).flatten(1)
        g = torch.sigmoid(F.linear(F.relu(F.linear(s, p["channel_attention.fc.0.weight"])),
                                   p["channel_attention.fc.2.weight"]))
        x = x * g.unsqueeze(2)
    for conv, bn, stride, pad in (("cnn_encoder.0", "cnn_encoder.1", 2, 3), ("cnn_encoder.4", "cnn_encoder.5", 2, 2)):
        x = F.conv1d(x, p[f"{conv}.weight"], None, stride, pad)
        x = F.batch_norm(x, bufs[f"{bn}.running_mean"], bufs[f"{bn}.running_var"], p[f"{bn}.weight"], p[f"{bn}.bias"],
                         training, 0.1, 1e-5)
        x = F.max_pool1d(F.relu(x), 3, 2, 1)
    x = x.permute(0, 2, 1)
    H = p["gru.weight_hh_l0"].shape[1]
    h0 = x.new_zeros(2 * layers, x.shape[0], H)
    out, _ = torch._VF.gru(x, h0, _gru_flat_weights(p, layers), True, layers, dropout if layers > 1 else 0.0,
                           training, True, True)
    last = out[:, -1, :]
    hid = F.dropout(F.relu(F.linear(last, p["classifier.0.weight"], p["classifier.0.bias"])), dropout, training)
    return F.linear(hid, p["classifier.3.weight"], p["classifier.3.bias"])


def make_state(C=6, num_classes=2, cnn_out=32, H=64, layers=2, seed=0):
    """Random-init parameters with the reference's shapes (values do not matter for timing)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: (torch.rand(*s, generator=g) - 0.5) * 0.2
    p = {"channel_attention.fc.0.weight": r(C // 4, C), "channel_attention.fc.2.weight": r(C, C // 4),
         "cnn_encoder.0.weight": r(16, C, 7), "cnn_encoder.1.weight": torch.ones(16), "cnn_encoder.1.bias": torch.zeros(16),
         "cnn_encoder.4.weight": r(cnn_out, 16, 5), "cnn_encoder.5.weight": torch.ones(cnn_out),
         "cnn_encoder.5.bias": torch.zeros(cnn_out)}
    for layer in range(layers):
        I = cnn_out if layer == 0 else 2 * H
        for suffix in ("", "_reverse"):
            p[f"gru.weight_ih_l{layer}{suffix}"] = r(3 * H, I)
            p[f"gru.weight_hh_l{layer}{suffix}"] = r(3 * H, H)
            p[f"gru.bias_ih_l{layer}{suffix}"] = r(3 * H)
            p[f"gru.bias_hh_l{layer}{suffix}"] = r(3 * H)
    p.update({"classifier.0.weight": r(64, 2 * H), "classifier.0.bias": r(64),
              "classifier.3.weight": r(num_classes, 64), "classifier.3.bias": r(num_classes)})
    bufs = {"cnn_encoder.1.running_mean": torch.zeros(16), "cnn_encoder.1.running_var": torch.ones(16),
            "cnn_encoder.5.running_mean": torch.zeros(cnn_out), "cnn_encoder.5.running_var": torch.ones(cnn_out)}
    return p, bufs


class CpuTrainStep:
    """zero_grad -> forward -> CE -> backward -> Adam.step (reference trainer.py:144-149) on CPU."""

    def __init__(self, p, bufs, lr=1e-3, weight_decay=1e-4, dropout=0.5, layers=2, attention=True):
        self.p = {k: v.clone().requires_grad_(v.numel() > 0) for k, v in p.items()}
        self.bufs = {k: v.clone() for k, v in bufs.items()}
        self.opt = torch.optim.Adam([v for v in self.p.values() if v.requires_grad], lr=lr, weight_decay=weight_decay)
        self.dropout, self.layers, self.attention = dropout, layers, attention

    def __call__(self, x, y):
        self.opt.zero_grad()
        logits = forward(self.p, self.bufs, x, training=True, dropout=self.dropout, layers=self.layers,
                         attention=self.attention)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        self.opt.step()
        return loss.item()                       # trainer.py:152 syncs on the loss every step


def time_train_steps(B=64, C=6, T=3840, steps=5, warmup=2, threads=None, seed=0):
    """windows/s of the CPU port on ``threads`` host threads (default: all)."""
    import os
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    p, bufs = make_state(C=C, seed=seed)
    step = CpuTrainStep(p, bufs)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, C, T, generator=g)
    y = torch.randint(0, 2, (B,), generator=g)
    for _ in range(warmup):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, y)
    dt = time.perf_counter() - t0
    return {"windows_per_s": B * steps / dt, "ms_per_step": 1e3 * dt / steps, "threads": threads,
            "steps": steps, "batch": B}
