"""CUDA-graph replay of the whole training step (reference train.py:368-401: forward, loss, backward, optimiser step).

One step of the hot path is a few hundred kernel launches issued from Python (`ops.py` -> ctypes -> `libdaftexprt_b200.so`); at
B = 32 the host needs longer to enqueue them than the B200 needs to run them, so the eager step is launch-bound.  Here the step
is captured ONCE per batch shape into a CUDA graph and replayed with a single launch:

    graph A (per batch shape): forward, loss, backward, gradients into the flat bucket            (all ranks, no collective)
    eager                    : ONE NCCL all-reduce of the flat bucket per OPTIMISER step           (world_size > 1 only)
    graph B (shape-free)     : clip_grad_norm_ (optional) + fused Adam over the flat buffers       (merged into graph A when
                                                                                                    world_size == 1 and no accumulation)

The loop shape of the reference is kept: `accumulation_steps` micro-batches (hparams.py:67, default 3) are accumulated before one
optimiser step (train.py:379,391-401), `grad_clip_thresh` is applied like `clip_grad_norm_` (train.py:399), and the learning rate
follows `update_learning_rate` (train.py:139-151) when `hparams` is given.  Differences that do not change the math: the gradient
mean over ranks is ONE all-reduce per optimiser step (the reference's DDP fires on every micro-step, train.py:293,391), and the
1 / accumulation_steps factor is applied to the gradients when they are added to the bucket instead of to the loss (the loss terms
returned are those of the micro-batch, undivided).

Everything that changes from step to step and used to be a kernel ARGUMENT lives in a 32-byte device block instead
(`dx_set_step_state`, include/daft_exprt_b200.h): the dropout seed epoch, the adversarial loss weight of this iteration
(loss.py:30-38), the learning rate and Adam's bias corrections.  The host writes the block (one 32-byte pinned H2D copy on the
replay stream) before each replay; the batch is copied into the graph's static input buffers.

Variable batch shapes: a graph is keyed on the padded shape (B, L_max, T_max).  `data.BucketedCollate` pads every batch up to a
small grid of bucket shapes, so a handful of graphs (LRU of `max_graphs`) covers a length distribution; `hits` / `misses` count
replays of cached graphs vs captures.  Multi-GPU: capture issues NO collective (the warm-up steps only prime kernels, allocator
pools and weight packs, and the weights are restored afterwards), so ranks may capture different shapes at different steps; the
only collective is the per-optimiser-step all-reduce, which every rank issues at the same micro-step count.
"""
import math
import struct
from collections import OrderedDict

import numpy as np
import torch

from . import cabi, ops

_MASK64 = (1 << 64) - 1


def reference_lr_schedule(hparams):
    """train.py:139-151 `update_learning_rate`: linear warm-up to max_learning_rate, then inverse-square-root decay."""
    lr0, lr1, warm = hparams.initial_learning_rate, hparams.max_learning_rate, hparams.warmup_steps

    def lr_of(iteration):
        if iteration < warm:
            return (lr1 - lr0) / warm * iteration + lr0
        return iteration ** -0.5 * lr1 / warm ** -0.5
    return lr_of


class StepState:
    """Host side of the per-step device block: a ring of pinned slots so the host may run ahead of the device."""
    SLOTS = 32

    def __init__(self, device):
        self.nbytes = int(cabi.load().dx_step_state_bytes())
        assert self.nbytes == 32, 'StepState layout changed: update graph.py'
        self.dev = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.host = torch.zeros(self.SLOTS, self.nbytes, dtype=torch.uint8).pin_memory()
        self.events = [None] * self.SLOTS
        self.count = 0

    def push(self, seed_epoch, w_adv, lr, bc1, bc2_sqrt):
        k = self.count % self.SLOTS
        self.count += 1
        if self.events[k] is not None:
            self.events[k].synchronize()   # the copy that last read this slot (SLOTS pushes ago) has run
        raw = struct.pack('<Qffffff', int(seed_epoch) & _MASK64, float(w_adv), float(lr), float(bc1), float(bc2_sqrt), 0.0, 0.0)
        self.host[k].numpy()[:] = np.frombuffer(raw, dtype=np.uint8)
        self.dev.copy_(self.host[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev

    def register(self, on=True):
        cabi.check(cabi.load().dx_set_step_state(self.dev.data_ptr() if on else None), 'dx_set_step_state')


class _Captured:
    __slots__ = ('inputs', 'targets', 'out', 'graph_a', 'launches')


class GraphedTrainStep:
    """step(inputs, targets, iteration) -> device tensor [8] with the weighted loss terms of this micro-batch.

    `model`, `criterion`, `sync`, `opt` are the eager objects (DaftExprt, DaftExprtLoss, FlatGradSync(mode='gather'), FlatAdam).
    Each call runs ONE micro-batch (forward, loss, backward); every `accumulation_steps`-th call also runs the gradient
    all-reduce, clip_grad_norm_ (opt.grad_clip_thresh) and the Adam step — `stepped` says whether the last call did.
    The returned tensor is a static buffer that the next replay of the same shape overwrites.
    """

    def __init__(self, model, criterion, sync, opt, lr_schedule=None, seed=None, max_graphs=8, warmup=2, accumulation_steps=1,
                 hparams=None, fused_exchange=None):
        self.model, self.criterion, self.sync, self.opt = model, criterion, sync, opt
        if lr_schedule is None and hparams is not None:
            lr_schedule = reference_lr_schedule(hparams)
        self.lr_schedule = lr_schedule            # callable(iteration) -> lr; None: opt.lr (e.g. set by the reference loop)
        self.max_graphs, self.warmup = max_graphs, warmup
        self.acc_steps = int(accumulation_steps)
        assert self.acc_steps >= 1
        self.device = sync.flat.device
        self.state = StepState(self.device)
        self.seed = torch.initial_seed() if seed is None else seed
        self.cache = OrderedDict()
        self.world = sync.world_size()
        assert sync.mode == 'gather', 'GraphedTrainStep needs FlatGradSync(mode="gather") (deferred weight-gradient reduction)'
        self.fused = fused_exchange   # ddp.FusedShardedAdam: reduce-scatter + Adam + all-gather in one kernel instead of all-reduce + graph B
        self.merged_adam = self.world == 1 and self.acc_steps == 1 and not opt.wants_clip()
        self.graph_b = None          # clip + Adam (+ zero the bucket when accumulating): independent of the batch shape
        self.launches_b = 0
        self.launches_replayed = 0   # kernels of libdaftexprt_b200.so executed through graph replays
        self.micro = 0               # micro-batches since the last optimiser step
        self.micro_total = 0         # dropout seed epoch
        self.hits = self.misses = 0
        self.stepped = False
        self.before_forward = self.before_backward = None   # optional host callbacks run while the step is captured (tools)
        if self.acc_steps > 1:
            self.sync.flat.zero_()

    # ------------------------------------------------------------------------------------------------------------------
    def _push_state(self, iteration):
        opt = self.opt
        self.micro_total += 1
        t = opt.step_count + 1    # number of the optimiser step this micro-batch belongs to
        lr = self.lr_schedule(iteration) if self.lr_schedule is not None else opt.lr
        bc1 = 1.0 - opt.betas[0] ** t
        bc2_sqrt = math.sqrt(1.0 - opt.betas[1] ** t)
        self.state.push(self.seed * 0x2545F4914F6CDD1D + self.micro_total, self.criterion.update_adversarial_weight(iteration), lr,
                        bc1, bc2_sqrt)

    def _body_backward(self, inputs, targets):
        self.opt.zero_grad()
        ops.set_wgrad_deferral(True)   # gather mode: no gradient is read before the bucket is filled, so reduce them all at once
        ops.use_grad_slots(self.acc_steps == 1)   # one micro-batch per optimiser step: the big weight gradients are written into the bucket
        try:
            if self.before_forward is not None:
                self.before_forward()
            out = self.criterion.forward_device(self.model(inputs), targets, 0)   # w_adv comes from the device block
            if self.before_backward is not None:
                self.before_backward()
            out[7].backward()
            ops.flush_wgrad()
        finally:
            ops.set_wgrad_deferral(False)
            ops.use_grad_slots(False)
        if self.acc_steps == 1:
            self.sync.gather()
        else:
            self.sync.accumulate(1.0 / self.acc_steps)
        return out

    def _body_adam(self):
        # step number / lr arguments are placeholders: the kernels read lr and the bias corrections from the device block
        self.opt.launch(1, 1.0 / self.world)
        if self.acc_steps > 1:
            self.sync.flat.zero_()     # the next optimiser step accumulates from zero

    def _all_reduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.sync.flat, op=torch.distributed.ReduceOp.SUM, group=self.sync.group)

    def _capture(self, inputs, targets):
        c = _Captured()
        # data.DeviceBatch (the inputs are views of one flat device buffer): keep that structure -> one copy per step
        c.inputs = inputs.clone() if hasattr(inputs, 'flat') else tuple(t.clone() for t in inputs)
        # targets that ARE input tensors (parse_batch returns the same objects, model.py:750) share the static buffer: one copy per step
        alias = self._alias(inputs, targets)
        c.targets = tuple(c.inputs[a] if a >= 0 else t.clone() for a, t in zip(alias, targets))
        lib = cabi.load()
        # warm-up on a side stream (lazy kernel attributes, allocator pools, weight packs).  NO collective here: ranks capture
        # independently (different shapes at different steps), so a warm-up all-reduce would pair with another rank's real one.
        # Parameters, moments and the gradient bucket are restored afterwards.
        backup = (self.opt.flat_p.clone(), self.opt.m.clone(), self.opt.v.clone(), self.sync.flat.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._push_state(0)
                self._body_backward(c.inputs, c.targets)
                self._body_adam()
                ops.invalidate_packed_weights()
            self.opt.flat_p.copy_(backup[0]); self.opt.m.copy_(backup[1]); self.opt.v.copy_(backup[2])
            self.sync.flat.copy_(backup[3])
            self.micro_total -= self.warmup
            ops.invalidate_packed_weights()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.zero_grad()
        l0 = lib.dx_launch_count()
        c.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph_a):
            c.out = self._body_backward(c.inputs, c.targets).detach()
            if self.merged_adam:
                self._body_adam()
        c.launches = int(lib.dx_launch_count() - l0)   # kernels of libdaftexprt_b200.so per replay
        if not self.merged_adam and self.graph_b is None and self.fused is None:
            l0 = lib.dx_launch_count()
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b):
                self._body_adam()
            self.launches_b = int(lib.dx_launch_count() - l0)
        ops.invalidate_packed_weights()
        return c

    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _alias(inputs, targets):
        ptrs = {t.data_ptr(): i for i, t in enumerate(inputs)}
        return tuple(ptrs.get(t.data_ptr(), -1) if t.numel() else -1 for t in targets)

    def _key(self, inputs, targets):
        return tuple(tuple(t.shape) for t in tuple(inputs) + tuple(targets)) + self._alias(inputs, targets)

    def step(self, inputs, targets, iteration):
        key = self._key(inputs, targets)
        self.state.register(True)
        try:
            c = self.cache.get(key)
            if c is None:
                self.misses += 1
                c = self._capture(inputs, targets)
                self.cache[key] = c
                while len(self.cache) > self.max_graphs:
                    self.cache.popitem(last=False)
            else:
                self.hits += 1
                self.cache.move_to_end(key)
            dsts, srcs, seen = [], [], set()
            pairs = zip(tuple(c.inputs) + tuple(c.targets), tuple(inputs) + tuple(targets))
            if hasattr(inputs, 'flat') and hasattr(c.inputs, 'flat') and all(a >= 0 for a in self._alias(inputs, targets)):
                if c.inputs.flat.data_ptr() != inputs.flat.data_ptr():
                    c.inputs.flat.copy_(inputs.flat, non_blocking=True)   # the whole batch: ONE device-to-device copy
                pairs = ()
            for dst, src in pairs:
                if dst.data_ptr() != src.data_ptr() and dst.data_ptr() not in seen:
                    seen.add(dst.data_ptr())
                    dsts.append(dst)
                    srcs.append(src)
            if dsts:
                torch._foreach_copy_(dsts, srcs)   # a few multi-tensor launches instead of one per tensor
            self._push_state(iteration)
            c.graph_a.replay()
            self.launches_replayed += c.launches
            self.micro += 1
            self.stepped = self.micro == self.acc_steps
            if self.stepped:
                self.micro = 0
                self.opt.step_count += 1
                if self.fused is not None:
                    self.fused.launch(self.opt.step_count)     # lr / bias corrections come from the device block
                    self.launches_replayed += 1
                    if self.acc_steps > 1:
                        self.sync.flat.zero_()
                elif not self.merged_adam:
                    self._all_reduce()
                    self.graph_b.replay()
                    self.launches_replayed += self.launches_b
        finally:
            self.state.register(False)
        if self.stepped:
            ops.invalidate_packed_weights()   # the replay moved the weights; packs cached by eager code are stale
        return c.out

    def static_batch(self, inputs, targets):
        """The graph's own input buffers for this shape (None before the first step): fill them directly to skip the copy."""
        c = self.cache.get(self._key(inputs, targets))
        return (c.inputs, c.targets) if c is not None else None
