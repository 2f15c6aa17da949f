"""Drop-in for the reference ``trainer.py``: ``Trainer(model, fold_output_dir, config)`` with
``.train(train_loader, val_loader)`` / ``.evaluate(loader, is_test, is_val)`` and
``EarlyStopping`` -- same config keys, log lines, ``best_model.pt`` and return values.

What changes is how a step executes (reference trainer.py:144-149): ``zero_grad -> forward ->
CrossEntropyLoss -> backward -> Adam.step`` is ONE C-ABI call (``mms_cnngru_train_step``) that
enqueues the whole kernel chain, replayed as a CUDA graph, with the loss accumulated on the
device (the reference syncs twice per step through ``loss.item()``, trainer.py:152-153).
"""
from __future__ import annotations

import ctypes as C
import time
from pathlib import Path

import numpy as np
import torch
from torch.optim.lr_scheduler import ReduceLROnPlateau

from . import _ext
from ._ext import CnnGruDesc, check, ptr, stream


class EarlyStopping:
    """reference trainer.py:12-39, bug-for-bug: the comparison treats the monitored
    score as higher-is-better although ``Trainer`` feeds it ``val_loss``, so ``best_model.pt`` is
    re-written whenever the validation loss does NOT improve on ``best_score + delta`` and the
    patience counter advances when it does.  ``fixed=True`` (off by default) monitors the loss
    the way the reference's comments intend."""

    def __init__(self, patience=7, delta=0, checkpoint_path='checkpoint.pt', verbose=False, log_func=None, fixed=False):
        self.patience, self.delta = patience, delta
        self.checkpoint_path = checkpoint_path
        self.verbose, self.log_func = verbose, log_func
        self.counter, self.best_score, self.early_stop = 0, None, False
        self.fixed = fixed

    def __call__(self, score, model):
        value = -score if self.fixed else score
        if self.best_score is None:
            self.best_score = value
            self.save_checkpoint(model)
            return
        if value < self.best_score + self.delta:
            self.counter += 1
            if self.verbose and self.log_func:
                self.log_func(f"EarlyStopping counter: {self.counter}/{self.patience}")
            self.early_stop = self.counter >= self.patience
        else:
            self.best_score = value
            self.save_checkpoint(model)
            self.counter = 0

    def save_checkpoint(self, model):
        # plain tensors (not views of the flat buffer) so the file equals the reference's layout
        torch.save({k: v.detach().clone() for k, v in model.state_dict().items()}, self.checkpoint_path)


class FlatAdam(torch.optim.Optimizer):
    """``torch.optim.Adam(model.parameters(), lr, weight_decay)`` (reference trainer.py:68) over
    the model's flat parameter buffer: one kernel for all 30 tensors.  It is a real
    ``Optimizer`` (``param_groups[0]['lr']`` is what ``ReduceLROnPlateau`` edits, trainer.py:72-77);
    the learning rate and the step count live on the device so the step can be graph-replayed."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in model.parameters()]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        flat = model.flat_parameters()
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.grads = torch.zeros_like(flat)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=flat.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=flat.device)
        self.scratch = torch.zeros(1, dtype=torch.int32, device=flat.device)
        self._lr_on_device = float(lr)

    def sync_lr(self):
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_on_device:
            self.lr_dev.fill_(lr)
            self._lr_on_device = lr

    @torch.no_grad()
    def step(self, closure=None):
        """Generic path: gather ``p.grad`` into the flat gradient buffer, then one Adam launch."""
        lib = _ext.lib()
        flat = self.model.flat_parameters()
        self.grads.zero_()
        for (off, n, _), p in zip(self.model._param_views, self.model.parameters()):
            if n and p.grad is not None:
                self.grads[off:off + n].copy_(p.grad.reshape(-1))
        self.flat_step(flat)

    def flat_step(self, flat=None):
        lib = _ext.lib()
        flat = self.model.flat_parameters() if flat is None else flat
        self.sync_lr()
        g = self.param_groups[0]
        check(lib.mms_adam_flat_step(ptr(flat), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), flat.numel(),
                                     ptr(self.lr_dev), g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'],
                                     ptr(self.step_dev), ptr(self.scratch), stream()))


class FusedTrainStep:
    """One training step = one ``mms_cnngru_train_step`` call, captured once per batch shape into
    a CUDA graph and replayed.  ``__call__(x, y)`` accepts device tensors or pinned host tensors
    (copied into the static graph inputs on the current stream)."""

    def __init__(self, model, optimizer: FlatAdam, batch: int, seq_len: int, use_graph: bool = True):
        self.lib = _ext.lib()
        self.model, self.opt = model, optimizer
        self.flat = model.flat_parameters()
        dev = self.flat.device
        self.x = torch.zeros(batch, model.in_channels, seq_len, dtype=torch.float32, device=dev)
        self.y = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.logits = torch.zeros(batch, model.num_classes, dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        d = CnnGruDesc()
        d.batch, d.in_channels, d.seq_len, d.num_classes = batch, model.in_channels, seq_len, model.num_classes
        d.cnn_out, d.hidden, d.layers = model.cnn_out_channels, model.gru_hidden_size, model.gru_num_layers
        d.training, d.attention, d.need_grad = 1, int(model.attention), 1
        d.dropout_p = float(model.dropout_p)
        d.rng_seed, d.rng_offset = model._rng_seed ^ 0x5DEECE66D, 0
        d.rng_offset_dev = optimizer.step_dev.data_ptr()      # dropout stream advances with the step count
        self.desc = d
        nbytes = self.lib.mms_cnngru_workspace_bytes(C.byref(d))
        if nbytes < 0:
            check(int(nbytes))
        self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        self.use_graph = use_graph
        self.graph = None
        self.calls = 0
        self.kernels_per_step = None

    def _enqueue(self):
        m, o, g = self.model, self.opt, self.opt.param_groups[0]
        check(self.lib.mms_cnngru_train_step(
            C.byref(self.desc), ptr(self.x), ptr(self.y), ptr(self.flat), ptr(o.grads), ptr(o.exp_avg), ptr(o.exp_avg_sq),
            ptr(m._bn_flat), ptr(m._nbt_flat), ptr(self.workspace), ptr(self.logits), ptr(self.loss), ptr(self.loss_sum),
            ptr(o.lr_dev), g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'], ptr(o.step_dev), ptr(o.scratch),
            stream()))

    def load(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)

    # ---- double-buffered input pipeline: the H2D copy of batch i+1 overlaps the step of batch i ----------
    def _pipeline(self):
        if getattr(self, "_copy_stream", None) is None:
            dev = self.x.device
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_x = [torch.empty_like(self.x) for _ in range(2)]
            self._stage_y = [torch.empty_like(self.y) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(2)]
            self._queued, self._next_slot = [], 0
        return self._copy_stream

    def prefetch(self, x, y):
        """Enqueue the host->device copy of an upcoming batch on the copy stream (staging slot)."""
        cs = self._pipeline()
        slot = self._next_slot
        cs.wait_event(self._consumed[slot])            # the step that last read this slot has copied it out
        if x.is_cuda:                                  # device-resident batch: produced on the current stream
            cs.wait_stream(torch.cuda.current_stream())
            x.record_stream(cs)
            y.record_stream(cs)
        with torch.cuda.stream(cs):
            self._stage_x[slot].copy_(x, non_blocking=True)
            self._stage_y[slot].copy_(y, non_blocking=True)
            self._ready[slot].record(cs)
        self._queued.append(slot)
        self._next_slot ^= 1

    def run_prefetched(self):
        """Run one step on the oldest prefetched batch (device-to-device hand-over, then the graph)."""
        slot = self._queued.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[slot])
        self.x.copy_(self._stage_x[slot], non_blocking=True)
        self.y.copy_(self._stage_y[slot], non_blocking=True)
        self._consumed[slot].record(cur)
        self.run()

    def run(self):
        """Enqueue one step on the static inputs (graph replay after the first two calls)."""
        if self.model.flat_parameters().data_ptr() != self.flat.data_ptr():
            raise _ext.MmsError("the model's parameter storage moved after FusedTrainStep was built")
        self.opt.sync_lr()
        self.calls += 1
        if not self.use_graph or self.calls == 1:
            self._enqueue()                       # first call eager: sets kernel attributes, warms caches
            return
        if self.graph is None:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._enqueue()
            self.graph = graph
        self.graph.replay()

    def post_loss(self):
        """Enqueue an asynchronous device->host copy of this step's loss into a pinned two-slot ring
        and return a ticket; ``read_loss(ticket)`` waits for exactly that copy.  Lets a caller read
        every step's loss (as reference trainer.py:152 does) one step late, so the GPU never idles
        while Python enqueues the next step."""
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_ev = [torch.cuda.Event() for _ in range(2)]
            self._loss_slot = 0
        k = self._loss_slot
        self._loss_host[k].copy_(self.loss, non_blocking=True)
        self._loss_ev[k].record(torch.cuda.current_stream())
        self._loss_slot ^= 1
        return k

    def read_loss(self, ticket) -> float:
        self._loss_ev[ticket].synchronize()
        return float(self._loss_host[ticket][0])

    def __call__(self, x, y):
        self.load(x, y)
        self.run()

    def last_loss(self) -> float:
        return float(self.loss.item())


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lib = _ext.lib()
        logits = logits.contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        check(lib.mms_cross_entropy(ptr(logits), ptr(labels.contiguous()), logits.shape[0], logits.shape[1], ptr(loss),
                                    ptr(dlogits), None, stream()))
        ctx.save_for_backward(dlogits)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None


class CrossEntropyLoss(torch.nn.Module):
    """``torch.nn.CrossEntropyLoss()`` (mean reduction, reference trainer.py:69) on the CUDA kernel."""

    def forward(self, logits, labels):
        return _CrossEntropy.apply(logits, labels)


class Trainer:
    """reference trainer.py:41-247."""

    def __init__(self, model, fold_output_dir: Path, config):
        self.model, self.fold_dir, self.config = model, Path(fold_output_dir), config
        self.fold_dir.mkdir(parents=True, exist_ok=True)
        self.log_file = self.fold_dir / 'training_log.txt'
        with open(self.log_file, 'w') as f:
            f.write(f"Training log for run starting at {time.strftime('%Y-%m-%d %H:%M:%S')}\n")
            f.write("=" * 50 + "\n")
        if not torch.cuda.is_available():
            raise _ext.MmsError("Trainer needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.model.to(self.device)

        t = config['trainer']
        self.epochs, self.learning_rate = t['epochs'], t['learning_rate']
        self.patience, self.weight_decay = t['early_stopping']['patience'], t['weight_decay']
        self.use_class_weights = t.get('use_class_weights', False)     # never enabled upstream
        self.use_graph = t.get('cuda_graph', True)

        self.optimizer = FlatAdam(self.model, lr=self.learning_rate, weight_decay=self.weight_decay)
        self.criterion = CrossEntropyLoss()
        self.scheduler = ReduceLROnPlateau(self.optimizer, mode='min', factor=0.1, patience=3)
        self.early_stopping = None
        if t['early_stopping']['enabled']:
            self.early_stopping = EarlyStopping(patience=self.patience, delta=t['early_stopping']['delta'],
                                                checkpoint_path=self.fold_dir / 'best_model.pt', verbose=True,
                                                log_func=self._log, fixed=t['early_stopping'].get('fixed', False))
        self._steps = {}
        self.total_start_time = time.time()
        self.windows_trained = 0

    def _log(self, message):
        print(message)
        with open(self.log_file, 'a') as f:
            f.write(message + '\n')

    def _fused(self, batch, seq_len):
        key = (batch, seq_len)
        if key not in self._steps:
            self._steps[key] = FusedTrainStep(self.model, self.optimizer, batch, seq_len, use_graph=self.use_graph)
        return self._steps[key]

    @staticmethod
    def _check_batch(batch):
        if isinstance(batch[0], (list, tuple)):
            raise NotImplementedError("HybridDataset inputs (reference void/dataset.py) are out of scope")

    def _loss_sum(self):
        return sum(float(s.loss_sum.item()) for s in self._steps.values())

    def train(self, train_loader, val_loader):
        best_val_acc = 0
        for epoch in range(self.epochs):
            t0 = time.time()
            self.model.train()
            for s in self._steps.values():
                s.loss_sum.zero_()
            # trainer.py:130-149 with a one-batch look-ahead: the H2D copy of the next batch runs on a copy
            # stream while the current step (one CUDA-graph replay) executes
            it = iter(train_loader)
            nxt = next(it, None)
            if nxt is not None:
                self._check_batch(nxt)
                self._fused(nxt[0].shape[0], nxt[0].shape[2]).prefetch(*nxt)
            while nxt is not None:
                inputs, labels = nxt
                step = self._fused(inputs.shape[0], inputs.shape[2])
                nxt = next(it, None)
                if nxt is not None:
                    self._check_batch(nxt)
                    self._fused(nxt[0].shape[0], nxt[0].shape[2]).prefetch(*nxt)
                step.run_prefetched()
                self.windows_trained += inputs.shape[0]
            train_loss = self._loss_sum()                  # one sync per epoch instead of two per step
            epoch_duration = time.time() - t0

            val_loss, val_acc, val_f1, val_preds, val_labels = self.evaluate(val_loader, is_val=True)
            self.scheduler.step(val_loss)
            best_val_acc = max(best_val_acc, val_acc)
            self._log(f"Epoch {epoch + 1}/{self.epochs} | "
                      f"耗时: {epoch_duration:.2f}s | "
                      f"训练损失: {train_loss / len(train_loader.dataset):.4f} | "
                      f"验证损失: {val_loss:.4f} | "
                      f"验证Acc: {val_acc:.4f} | "
                      f"验证F1: {val_f1:.4f}")
            if self.early_stopping:
                self.early_stopping(val_loss, self.model)
                if self.early_stopping.early_stop:
                    self._log("触发早停")
                    break
        if self.early_stopping and self.early_stopping.early_stop:
            self._log(f"加载性能最佳的模型权重从: {self.early_stopping.checkpoint_path}")
            self.model.load_state_dict(torch.load(self.early_stopping.checkpoint_path, weights_only=True))
        self._log(f"--- 训练完成 --- 总训练时长: {time.time() - self.total_start_time:.2f}秒")

    def evaluate(self, data_loader, is_test=False, is_val=False):
        from sklearn.metrics import accuracy_score, f1_score
        self.model.eval()
        loss_acc = torch.zeros(1, dtype=torch.float64, device=self.device)
        preds, labels_all = [], []
        lib = _ext.lib()
        scratch = torch.empty(1, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for inputs, labels in data_loader:
                inputs = inputs.to(self.device, non_blocking=True)
                labels = labels.to(self.device, non_blocking=True)
                outputs = self.model(inputs)
                check(lib.mms_cross_entropy(ptr(outputs), ptr(labels), outputs.shape[0], outputs.shape[1], ptr(scratch),
                                            None, ptr(loss_acc), stream()))
                preds.append(torch.argmax(outputs, dim=1))          # argmax(softmax) == argmax(logits), trainer.py:224-225
                labels_all.append(labels)
        all_preds = torch.cat(preds).cpu().numpy()
        all_labels = torch.cat(labels_all).cpu().numpy()
        loss = float(loss_acc.item()) / len(data_loader.dataset)
        acc = accuracy_score(all_labels, all_preds)
        f1 = f1_score(all_labels, all_preds, average='weighted')
        if is_test:
            self.plot_confusion_matrix(all_labels, all_preds, filename="test_confusion_matrix.png")
            self._log(f"\n--- 最终测试结果 (模型原始输出) ---")
            self._log(f"测试损失: {loss:.4f} | 测试Acc: {acc:.4f} | 测试F1: {f1:.4f}")
            return loss, acc, f1
        if is_val:
            return loss, acc, f1, list(all_preds), list(all_labels)
        return loss, acc, f1

    def plot_confusion_matrix(self, true_labels, pred_labels, filename="confusion_matrix.png"):
        """Plotting is out of scope; like the reference's own try/except
        (trainer.py:250-273) a missing matplotlib only produces a log line."""
        try:
            from sklearn.metrics import confusion_matrix
            import matplotlib.pyplot as plt
            cm = confusion_matrix(true_labels, pred_labels)
            fig = plt.figure(figsize=(8, 6))
            plt.imshow(cm, cmap='Blues')
            for (i, j), v in np.ndenumerate(cm):
                plt.text(j, i, str(v), ha='center', va='center')
            plt.xlabel('Predicted Label')
            plt.ylabel('True Label')
            plt.title('Confusion Matrix')
            cm_path = self.fold_dir / filename
            plt.savefig(cm_path)
            plt.close(fig)
            self._log(f"混淆矩阵已保存至: {cm_path}")
        except Exception as e:
            self._log(f"保存混淆矩阵失败: {str(e)}")
