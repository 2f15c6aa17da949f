"""Intra-fold data parallelism (BASELINE.json configs[4]): the global batch is split over the
ranks of a process group; every rank holds a replica of the model.

The reference computes BatchNorm statistics over the WHOLE batch on one device (models.py:47,51), so a
faithful data-parallel step needs SyncBN: the float64 batch sums (forward) and the two BN reductions
(backward) are sum-all-reduced between the phases of ``mms_cnngru_{forward,backward}_phase``; the loss is
the global mean (``mms_cross_entropy_partial``); the flat gradient buffer is sum-all-reduced ONCE per
step and the fused Adam kernel then runs on every rank.  All messages are tiny (<= 0.5 MB): latency-bound
on NVLink 5 / NVSwitch, which is why the gradient is one flat buffer and one collective.

The step is written as a generator that yields the tensors to reduce, so the same code runs under
``torch.distributed`` (NCCL) and in the single-process emulation the parity tests use.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _ext
from ._ext import CnnGruDesc, check, ptr, stream
from .trainer import FlatAdam


class DataParallelTrainStep:
    """One rank's share of a data-parallel training step.  ``local_batch`` rows of a ``global_batch`` batch."""

    def __init__(self, model, optimizer: FlatAdam, local_batch: int, global_batch: int, seq_len: int, rank: int = 0, group=None):
        self.lib = _ext.lib()
        self.model, self.opt, self.group = model, optimizer, group
        self.flat = model.flat_parameters()
        dev = self.flat.device
        d = CnnGruDesc()
        d.batch, d.in_channels, d.seq_len, d.num_classes = local_batch, model.in_channels, seq_len, model.num_classes
        d.cnn_out, d.hidden, d.layers = model.cnn_out_channels, model.gru_hidden_size, model.gru_num_layers
        d.training, d.attention, d.need_grad = 1, int(model.attention), 1
        d.dropout_p = float(model.dropout_p)
        d.rng_seed, d.rng_offset = (model._rng_seed ^ (0x9E3779B9 * (rank + 1))) & 0xFFFFFFFFFFFFFFFF, 0
        d.rng_offset_dev = optimizer.step_dev.data_ptr()
        d.global_batch = global_batch
        self.desc = d
        self.local_batch, self.global_batch = local_batch, global_batch
        nbytes = self.lib.mms_cnngru_workspace_bytes(C.byref(d))
        if nbytes < 0:
            check(int(nbytes))
        self.workspace = torch.zeros(int(nbytes), dtype=torch.uint8, device=dev)
        offs = (C.c_int64 * 4)()
        cnts = (C.c_int64 * 4)()
        check(self.lib.mms_cnngru_sync_offsets(C.byref(d), offs, cnts))
        view = lambda i: self.workspace[offs[i]:offs[i] + 8 * cnts[i]].view(torch.float64)
        self.stats1, self.stats2, self.red1, self.red2 = view(0), view(1), view(2), view(3)
        self.logits = torch.zeros(local_batch, model.num_classes, dtype=torch.float32, device=dev)
        self.dlogits = torch.zeros_like(self.logits)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)        # this rank's share of the global mean

    def _fwd(self, phases, x):
        m = self.model
        check(self.lib.mms_cnngru_forward_phase(C.byref(self.desc), phases, ptr(x), ptr(self.flat), ptr(m._bn_flat), ptr(m._nbt_flat),
                                                ptr(self.workspace), ptr(self.logits), stream()))

    def _bwd(self, phases, x):
        m, o = self.model, self.opt
        check(self.lib.mms_cnngru_backward_phase(C.byref(self.desc), phases, ptr(x), ptr(self.flat), ptr(m._bn_flat),
                                                 ptr(self.workspace), ptr(self.dlogits), ptr(o.grads), stream()))

    def phases(self, x, y):
        """Generator over the step: yields each tensor that must be sum-reduced over the ranks before it resumes."""
        x = x.contiguous().float()
        y = y.contiguous()
        self.opt.grads.zero_()                                   # trainer.py:144
        self._fwd(1, x)
        yield self.stats1                                        # SyncBN, stage 1: (sum, sum of squares) per channel
        self._fwd(2, x)
        yield self.stats2
        self._fwd(4, x)                                          # trainer.py:146
        check(self.lib.mms_cross_entropy_partial(ptr(self.logits), ptr(y), self.local_batch, self.model.num_classes,
                                                 self.global_batch, ptr(self.loss), ptr(self.dlogits), None, stream()))   # :147
        self._bwd(1, x)                                          # trainer.py:148
        yield self.red2                                          # SyncBN backward, stage 2: (sum dy, sum dy*xhat)
        self._bwd(2, x)
        yield self.red1
        self._bwd(4, x)
        yield self.opt.grads                                     # ONE flat gradient all-reduce (~0.5 MB)
        self.opt.flat_step(self.flat)                            # trainer.py:149

    def __call__(self, x, y):
        import torch.distributed as dist
        for t in self.phases(x, y):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def global_loss(self) -> float:
        """Mean loss over the global batch of the last step (one extra scalar all-reduce; for logging)."""
        import torch.distributed as dist
        t = self.loss.clone()
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())


def emulate_ranks(steps, batches):
    """Run several ranks' ``DataParallelTrainStep`` in lockstep inside ONE process (no collectives):
    at every sync point the yielded tensors are summed and written back to every rank.  This is the
    single-GPU stand-in for NCCL that the parity tests use."""
    gens = [s.phases(x, y) for s, (x, y) in zip(steps, batches)]
    while True:
        tensors = []
        for g in gens:
            try:
                tensors.append(next(g))
            except StopIteration:
                pass
        if not tensors:                      # every rank has run its Adam update
            break
        assert len(tensors) == len(gens), "ranks fell out of lockstep"
        total = torch.stack(tensors).sum(dim=0)
        for t in tensors:
            t.copy_(total)
