"""Data-parallel plumbing for the hot path: ONE flat fp32 gradient buffer, ONE all-reduce per optimiser step.

The reference wraps the model in stock DistributedDataParallel (train.py:293): ~3 buckets of <= 25 MB, fired on every
micro-step even under gradient accumulation (no `no_sync()`, train.py:391).  Utterances are independent, so the only
exchange on this path is the mean of the 14.7 M-element gradient (58.9 MB fp32).  Here every `p.grad` is a view into one
flat buffer, so the exchange is a single NCCL all-reduce over NVLink 5 / NVSwitch (in-switch reduction when NCCL picks NVLS)
issued once per optimiser step, and the optimiser is one fused kernel over the same flat layout (`dx_adam_step`).
`torch.distributed` is used for process-group plumbing only; one process per GPU.
"""
import torch
import torch.distributed as dist

ALIGN = 64   # elements (256 bytes of fp32)


class FlatGradSync:
    """One flat fp32 gradient bucket for all parameters, averaged across ranks with one collective.

    mode='alias'  : every `p.grad` is a view of the flat buffer; autograd accumulates into it in place (one small add kernel
                    per parameter per backward) — the layout stock optimisers can also consume.
    mode='gather' : gradients are left to autograd (`p.grad = None` before backward, so AccumulateGrad just keeps the tensors
                    the backward kernels produced: no add kernels, no zeroing pass) and are packed into the flat buffer by ONE
                    multi-tensor copy before the all-reduce.  Used by bench.py.
    """

    def __init__(self, params, process_group=None, mode='alias'):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        assert mode in ('alias', 'gather')
        self.mode = mode
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        # every slot starts on a 256-byte boundary: the kernels use 16-byte vector loads on parameters and gradients
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.padded_numel = off
        self.flat = torch.zeros(self.padded_numel, device=dev, dtype=dt)
        self.group = process_group
        self.views = [self.flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        if mode == 'alias':
            for p, v in zip(self.params, self.views):
                p.grad = v   # autograd accumulates in place into existing .grad
        self._register_slots()

    def _register_slots(self):
        """mode 'gather' on CUDA: tell the weight-gradient GEMMs where each parameter's gradient lives in the bucket (ops.grad_slot);
        used while `ops.use_grad_slots(True)` (one micro-batch per optimiser step: graph.GraphedTrainStep / bench switch it on)."""
        if self.mode == 'gather' and self.flat.is_cuda:
            from . import ops
            ops.register_grad_slots({id(p): v for p, v in zip(self.params, self.views)})

    def rebind(self, new_flat):
        """Move the bucket into another allocation of the same size (e.g. a symmetric-memory buffer mapped on every rank)."""
        assert new_flat.numel() == self.padded_numel and new_flat.dtype == self.flat.dtype
        new_flat.copy_(self.flat)
        self.flat = new_flat
        self.views = [self.flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        if self.mode == 'alias':
            for p, v in zip(self.params, self.views):
                p.grad = v
        self._register_slots()

    def zero_grad(self):
        if self.mode == 'alias':
            self.flat.zero_()
        else:
            for p in self.params:
                p.grad = None

    def gather(self):
        """mode='gather': pack the per-parameter gradients into the flat bucket (one multi-tensor copy)."""
        if self.mode != 'gather':
            return
        views, grads = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():   # (gradients written straight into their slot need no copy: ops.grad_slot)
                views.append(v)
                grads.append(p.grad)
        if views:
            torch._foreach_copy_(views, grads)

    def accumulate(self, alpha=1.0):
        """mode='gather', gradient accumulation (reference train.py:379,391: loss / accumulation_steps, backward per micro-batch):
        flat += alpha * grads with ONE multi-tensor add, then drop the per-parameter gradients so the next micro-batch's
        backward starts clean.  The caller zeroes `flat` once per optimiser step."""
        assert self.mode == 'gather'
        views, grads = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                views.append(v)
                grads.append(p.grad)
        if views:
            torch._foreach_add_(views, grads, alpha=alpha)
        for p in self.params:
            p.grad = None

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce_mean(self, async_op=False):
        """Average gradients over ranks (DDP semantics).  No-op for a single process."""
        self.gather()
        ws = self.world_size()
        if ws == 1:
            return None
        self.flat.div_(ws)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def grads_attached(self):
        """mode='alias': True while every p.grad still aliases the flat buffer (zero_grad(set_to_none=True) breaks it)."""
        base = self.flat.data_ptr()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != base + off * self.flat.element_size():
                return False
        return True


def broadcast_parameters(module, src=0, process_group=None):
    """Rank-`src` parameters to every rank (what DDP's constructor does, train.py:293)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
    from . import ops
    ops.invalidate_packed_weights()   # p.data writes do not bump tensor versions: cached weight packs are stale on non-src ranks


class FlatAdam:
    """torch.optim.Adam semantics (reference train.py:299-301: betas (0.9, 0.98), eps 1e-9, weight_decay 1e-6) as ONE fused
    kernel over flat parameter / gradient / moment buffers.  Parameters are re-pointed at views of the flat buffer.

    The parts of the torch.optim interface the reference's loop touches are kept: `param_groups` (train.py:316,405-408 read and
    write `param_group['lr']`), `state_dict()` / `load_state_dict()` in torch.optim.Adam's own format (save/load_checkpoint,
    train.py:77,122-128 — a checkpoint written by the reference's Adam loads here and vice versa), `zero_grad()`, `step()`.
    `grad_clip_thresh` reproduces `clip_grad_norm_` (train.py:399) on the device; `last_grad_norm()` reads the norm back."""

    def __init__(self, params, sync: FlatGradSync, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6,
                 grad_clip_thresh=float('inf'), track_grad_norm=False):
        self.sync = sync
        self.params = sync.params
        assert [id(p) for p in self.params] == [id(p) for p in params if p.requires_grad]
        self.param_groups = [dict(params=self.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                  amsgrad=False, maximize=False)]
        self.grad_clip_thresh = float(grad_clip_thresh)
        self.track_grad_norm = bool(track_grad_norm)
        self.flat_p = torch.zeros_like(sync.flat)
        with torch.no_grad():
            for p, off in zip(self.params, sync.offsets):
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.grad_stats = torch.zeros(4, device=self.flat_p.device, dtype=torch.float32)   # {norm, clip coefficient, scratch}
        self.step_count = 0

    # the scalar hyper-parameters live in param_groups[0] (what the reference's loop edits); these are views of it
    lr = property(lambda self: self.param_groups[0]['lr'], lambda self, v: self.param_groups[0].__setitem__('lr', v))
    betas = property(lambda self: self.param_groups[0]['betas'], lambda self, v: self.param_groups[0].__setitem__('betas', tuple(v)))
    eps = property(lambda self: self.param_groups[0]['eps'], lambda self, v: self.param_groups[0].__setitem__('eps', v))
    weight_decay = property(lambda self: self.param_groups[0]['weight_decay'],
                            lambda self, v: self.param_groups[0].__setitem__('weight_decay', v))

    def rebind(self, new_flat_p):
        """Move the flat parameter buffer (and every `p.data` view of it) into another allocation of the same size."""
        assert new_flat_p.numel() == self.flat_p.numel()
        with torch.no_grad():
            new_flat_p.copy_(self.flat_p)
            self.flat_p = new_flat_p
            for p, off in zip(self.params, self.sync.offsets):
                p.data = self.flat_p[off:off + p.numel()].view_as(p)
        from . import ops
        ops.invalidate_packed_weights()

    def wants_clip(self):
        return self.track_grad_norm or self.grad_clip_thresh != float('inf')

    def launch(self, step, grad_scale, stream=None):
        """Issue (clip +) Adam for optimiser step number `step` on the current stream; no host-side state changes."""
        from . import ops
        st = ops._st() if stream is None else stream
        clip = None
        if self.wants_clip():
            ops._call('dx_grad_norm_clip', self.sync.flat.data_ptr(), self.sync.flat.numel(), float(grad_scale),
                      float(self.grad_clip_thresh), self.grad_stats.data_ptr(), st)
            clip = self.grad_stats.data_ptr()
        ops._call('dx_adam_step', self.flat_p.data_ptr(), self.sync.flat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                  self.flat_p.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                  float(self.weight_decay), int(step), float(grad_scale), clip, st)

    def step(self, grad_scale=1.0):
        from . import ops
        self.step_count += 1
        self.launch(self.step_count, grad_scale)
        ops.invalidate_packed_weights()   # raw-pointer update does not bump tensor versions

    def last_grad_norm(self):
        """Gradient norm of the last step (device->host read; needs grad_clip_thresh < inf or track_grad_norm)."""
        return float(self.grad_stats[0].item())

    def zero_grad(self, set_to_none=True):
        self.sync.zero_grad()

    # -- torch.optim.Adam checkpoint format ---------------------------------------------------------------------------------
    def state_dict(self):
        """torch.optim.Adam's format.  With ddp.FusedShardedAdam attached every rank keeps the moments of its own shard only: they are
        reassembled here (a COLLECTIVE call then — every rank must call it, like the reference's barrier around checkpoints)."""
        state = {}
        fused = getattr(self, '_fused', None)
        m, v = fused.full_moments() if fused is not None else (self.m, self.v)
        if self.step_count > 0:
            for i, (p, off) in enumerate(zip(self.params, self.sync.offsets)):
                n = p.numel()
                state[i] = {'step': torch.tensor(float(self.step_count)),
                            'exp_avg': m[off:off + n].view_as(p).clone(),
                            'exp_avg_sq': v[off:off + n].view_as(p).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        group['params'] = list(range(len(self.params)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        groups = sd['param_groups']
        if len(groups) != 1 or len(groups[0]['params']) != len(self.params):
            raise ValueError('FlatAdam.load_state_dict: expected one param group covering every parameter')
        for k, v in groups[0].items():
            if k != 'params' and k in self.param_groups[0]:
                self.param_groups[0][k] = tuple(v) if k == 'betas' else v
        self.m.zero_()
        self.v.zero_()
        steps = set()
        with torch.no_grad():
            for i, st in sd['state'].items():
                off, p = self.sync.offsets[int(i)], self.params[int(i)]
                n = p.numel()
                self.m[off:off + n].copy_(st['exp_avg'].reshape(-1))
                self.v[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
                steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError('FlatAdam.load_state_dict: parameters with different step counts are not supported')
        self.step_count = steps.pop() if steps else 0
        fused = getattr(self, '_fused', None)
        if fused is not None:   # sharded moments: keep the own shard only
            for buf in (self.m, self.v):
                buf[:fused.begin].zero_()
                buf[fused.begin + fused.n:].zero_()


class FusedShardedAdam:
    """Gradient exchange fused with the optimiser over NVLink / NVSwitch peer memory (multi-GPU, one process per GPU).

    Replaces `all-reduce of the flat bucket -> Adam over the whole bucket` (the reference: DDP all-reduce + torch.optim.Adam,
    train.py:293,391,401) by ONE kernel per rank (`dx_fused_reduce_adam`): reduce-scatter of this rank's shard of the gradient
    (`multimem.ld_reduce`: the NVSwitch adds the N copies; peer loads when the buffers have no multicast mapping), Adam on the shard,
    all-gather of the updated parameters (`multimem.st` / peer stores), bracketed by two device-side cross-GPU barriers.  The flat
    gradient and parameter buffers move into symmetric memory (`torch.distributed._symmetric_memory`); the Adam moments are only
    maintained for the own shard (ZeRO-1 style; `full_moments()` reassembles them for a checkpoint).

    Construction is a COLLECTIVE (every rank must call it); it raises if symmetric memory cannot be set up on this system — callers
    keep the NCCL all-reduce path then (`bench.py` does).  Gradient clipping needs the global norm and is not supported here."""

    def __init__(self, sync: FlatGradSync, opt: 'FlatAdam', group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        assert dist.is_available() and dist.is_initialized(), 'FusedShardedAdam needs an initialised process group'
        assert not opt.wants_clip(), 'FusedShardedAdam: clip_grad_norm_ needs the global gradient norm (use the all-reduce path)'
        self.sync, self.opt = sync, opt
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        assert self.world <= 16
        dev, P = sync.flat.device, sync.padded_numel
        g = symm.empty(P, dtype=torch.float32, device=dev)
        p = symm.empty(P, dtype=torch.float32, device=dev)
        self.hg = symm.rendezvous(g, self.group)
        self.hp = symm.rendezvous(p, self.group)
        sync.rebind(g)
        opt.rebind(p)
        shard = (P + self.world - 1) // self.world
        shard = (shard + ALIGN - 1) // ALIGN * ALIGN
        self.begin = min(self.rank * shard, P)
        self.n = min(shard, P - self.begin)
        self.g_mc = int(getattr(self.hg, 'multicast_ptr', 0) or 0)
        self.p_mc = int(getattr(self.hp, 'multicast_ptr', 0) or 0)
        if not (self.g_mc and self.p_mc):
            self.g_mc = self.p_mc = 0
        self.g_peers = (ctypes.c_uint64 * self.world)(*[int(x) for x in self.hg.buffer_ptrs])
        self.p_peers = (ctypes.c_uint64 * self.world)(*[int(x) for x in self.hp.buffer_ptrs])
        self.mode = 'multimem (in-switch reduction)' if self.g_mc else 'peer loads / stores'
        opt._fused = self
        for buf in (opt.m, opt.v):            # moments of the other shards are not maintained on this rank
            buf[:self.begin].zero_()
            buf[self.begin + self.n:].zero_()
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def launch(self, step):
        """barrier -> reduce-scatter + Adam + all-gather -> barrier, all on the current stream; `step` = optimiser step number
        (ignored while a per-step device block is registered: lr and the bias corrections come from it)."""
        import ctypes
        from . import ops
        opt = self.opt
        self.hg.barrier(channel=0)          # every rank has finished writing its gradients
        ops._call('dx_fused_reduce_adam', self.g_mc or None, ctypes.addressof(self.g_peers), self.p_mc or None, ctypes.addressof(self.p_peers),
                  opt.flat_p.data_ptr(), opt.m.data_ptr(), opt.v.data_ptr(), self.begin, self.n, self.world, float(opt.lr),
                  float(opt.betas[0]), float(opt.betas[1]), float(opt.eps), float(opt.weight_decay), int(step), 1.0 / self.world, ops._st())
        self.hp.barrier(channel=1)          # every rank's parameter shard has landed everywhere

    def step(self):
        from . import ops
        self.sync.gather()
        self.opt.step_count += 1
        self.launch(self.opt.step_count)
        ops.invalidate_packed_weights()

    def full_moments(self):
        """(exp_avg, exp_avg_sq) over the whole flat layout: every rank holds its own shard only (the rest is zero) -> sum over ranks."""
        m, v = self.opt.m.clone(), self.opt.v.clone()
        dist.all_reduce(m, group=self.group)
        dist.all_reduce(v, group=self.group)
        return m, v
