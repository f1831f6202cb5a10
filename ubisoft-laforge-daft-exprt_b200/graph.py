"""CUDA-graph replay of the whole training step (reference train.py:368-401: forward, loss, backward, optimiser step).

One step of the hot path is ~640 kernel launches issued from Python (`ops.py` -> ctypes -> `libdaftexprt_b200.so`); at B = 32 the
host needs ~17.7 ms to enqueue them while the B200 needs ~17 ms to run them, so the eager step is launch-bound.  Here the step
is captured ONCE per batch shape into a CUDA graph and replayed with a single launch:

    graph A: zero grads, forward, loss, backward, pack gradients into the flat bucket        (all ranks, no collective)
    eager  : ONE NCCL all-reduce of the flat bucket                                           (world_size > 1 only)
    graph B: fused Adam over the flat buffers (gradient mean folded into grad_scale)          (merged into graph A at world_size 1)

Everything that changes from step to step and used to be a kernel ARGUMENT lives in a 32-byte device block instead
(`dx_set_step_state`, include/daft_exprt_b200.h): the dropout seed epoch, the adversarial loss weight of this iteration
(loss.py:30-38), the learning rate (train.py:139-151 schedules it per iteration) and Adam's bias corrections.  The host writes
the block (one 32-byte pinned H2D copy on the replay stream) before each replay; the batch is copied into the graph's static
input buffers.  Batches whose padded shape (B, L_max, T_max) was not seen before are captured on first use (LRU of
`max_graphs` graphs); pad batches to a few bucket shapes in the collate function to keep that set small.
"""
import math
import struct
from collections import OrderedDict

import numpy as np
import torch

from . import cabi, ops

_MASK64 = (1 << 64) - 1


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
    __slots__ = ('inputs', 'targets', 'out', 'graph_a', 'graph_b', 'launches')


class GraphedTrainStep:
    """step(inputs, targets, iteration) -> device tensor [8] with the weighted loss terms (same as `forward_device`).

    `model`, `criterion`, `sync`, `opt` are the eager objects (DaftExprt, DaftExprtLoss, FlatGradSync, FlatAdam); the captured
    step is exactly `zero_grad; criterion.forward_device(model(inputs), targets, it)[7].backward(); sync; opt.step()`.
    The returned tensor is a static buffer that the next replay overwrites.
    """

    def __init__(self, model, criterion, sync, opt, lr_schedule=None, seed=None, max_graphs=4, warmup=2):
        self.model, self.criterion, self.sync, self.opt = model, criterion, sync, opt
        self.lr_schedule = lr_schedule            # callable(iteration) -> lr, default: opt.lr
        self.max_graphs, self.warmup = max_graphs, warmup
        self.device = sync.flat.device
        self.state = StepState(self.device)
        self.seed = torch.initial_seed() if seed is None else seed
        self.cache = OrderedDict()
        self.world = sync.world_size()
        assert sync.mode == 'gather', 'GraphedTrainStep needs FlatGradSync(mode="gather") (deferred weight-gradient reduction)'
        self.launches_replayed = 0   # kernels of libdaftexprt_b200.so executed through graph replays

    # ------------------------------------------------------------------------------------------------------------------
    def _push_state(self, iteration):
        opt = self.opt
        opt.step_count += 1
        t = opt.step_count
        lr = self.lr_schedule(iteration) if self.lr_schedule is not None else opt.lr
        bc1 = 1.0 - opt.betas[0] ** t
        bc2_sqrt = math.sqrt(1.0 - opt.betas[1] ** t)
        self.state.push(self.seed * 0x2545F4914F6CDD1D + t, self.criterion.update_adversarial_weight(iteration), lr, bc1, bc2_sqrt)

    def _body_backward(self, inputs, targets):
        self.opt.zero_grad()
        ops.set_wgrad_deferral(True)   # gather mode: no gradient is read before sync.gather(), so reduce them all at once
        try:
            out = self.criterion.forward_device(self.model(inputs), targets, 0)   # w_adv comes from the device block
            out[7].backward()
            ops.flush_wgrad()
        finally:
            ops.set_wgrad_deferral(False)
        self.sync.gather()
        return out

    def _body_adam(self):
        # step number / lr arguments are placeholders: the kernel reads lr and the bias corrections from the device block
        ops._call('dx_adam_step', self.opt.flat_p.data_ptr(), self.sync.flat.data_ptr(), self.opt.m.data_ptr(),
                  self.opt.v.data_ptr(), self.opt.flat_p.numel(), float(self.opt.lr), float(self.opt.betas[0]),
                  float(self.opt.betas[1]), float(self.opt.eps), float(self.opt.weight_decay), 1, 1.0 / self.world, ops._st())

    def _all_reduce(self):
        if self.world > 1:
            torch.distributed.all_reduce(self.sync.flat, op=torch.distributed.ReduceOp.SUM, group=self.sync.group)

    def _capture(self, inputs, targets):
        c = _Captured()
        c.inputs = tuple(t.clone() for t in inputs)
        # targets that ARE input tensors (parse_batch returns the same objects, model.py:750) share the static buffer: one copy per step
        alias = self._alias(inputs, targets)
        c.targets = tuple(c.inputs[a] if a >= 0 else t.clone() for a, t in zip(alias, targets))
        lib = cabi.load()
        # warm-up on a side stream (lazy kernel attributes, allocator pools, weight packs), parameters restored afterwards
        backup = (self.opt.flat_p.clone(), self.opt.m.clone(), self.opt.v.clone(), self.opt.step_count)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._push_state(0)
                self._body_backward(c.inputs, c.targets)
                self._all_reduce()
                self._body_adam()
                ops.invalidate_packed_weights()
            self.opt.flat_p.copy_(backup[0]); self.opt.m.copy_(backup[1]); self.opt.v.copy_(backup[2])
            self.opt.step_count = backup[3]
            ops.invalidate_packed_weights()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.zero_grad()
        l0 = lib.dx_launch_count()
        c.graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph_a):
            c.out = self._body_backward(c.inputs, c.targets).detach()
            if self.world == 1:
                self._body_adam()
        c.graph_b = None
        if self.world > 1:
            c.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(c.graph_b):
                self._body_adam()
        c.launches = int(lib.dx_launch_count() - l0)   # kernels of libdaftexprt_b200.so per replay
        ops.invalidate_packed_weights()
        return c

    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _alias(inputs, targets):
        ptrs = {t.data_ptr(): i for i, t in enumerate(inputs)}
        return tuple(ptrs.get(t.data_ptr(), -1) if t.numel() else -1 for t in targets)

    def step(self, inputs, targets, iteration):
        key = tuple(tuple(t.shape) for t in tuple(inputs) + tuple(targets)) + self._alias(inputs, targets)
        self.state.register(True)
        try:
            c = self.cache.get(key)
            if c is None:
                c = self._capture(inputs, targets)
                self.cache[key] = c
                while len(self.cache) > self.max_graphs:
                    self.cache.popitem(last=False)
            else:
                self.cache.move_to_end(key)
            dsts, srcs, seen = [], [], set()
            for dst, src in zip(c.inputs + c.targets, tuple(inputs) + tuple(targets)):
                if dst.data_ptr() != src.data_ptr() and dst.data_ptr() not in seen:
                    seen.add(dst.data_ptr())
                    dsts.append(dst)
                    srcs.append(src)
            if dsts:
                torch._foreach_copy_(dsts, srcs)   # a few multi-tensor launches instead of one per tensor
            self._push_state(iteration)
            c.graph_a.replay()
            if c.graph_b is not None:
                self._all_reduce()
                c.graph_b.replay()
        finally:
            self.state.register(False)
        ops.invalidate_packed_weights()   # the replay moved the weights; packs cached by eager code are stale
        self.launches_replayed += c.launches
        return c.out

    def static_batch(self, inputs, targets):
        """The graph's own input buffers for this shape (None before the first step): fill them directly to skip the copy."""
        c = self.cache.get(tuple(tuple(t.shape) for t in tuple(inputs) + tuple(targets)) + self._alias(inputs, targets))
        return (c.inputs, c.targets) if c is not None else None
