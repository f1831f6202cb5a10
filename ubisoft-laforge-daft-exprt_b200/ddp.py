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
            else:
                views.append(v)
                grads.append(p.grad)
        torch._foreach_copy_(views, grads)

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


class FlatAdam:
    """torch.optim.Adam semantics (reference train.py:299-301: betas (0.9, 0.98), eps 1e-9, weight_decay 1e-6) as ONE fused
    kernel over flat parameter / gradient / moment buffers.  Parameters are re-pointed at views of the flat buffer."""

    def __init__(self, params, sync: FlatGradSync, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6):
        self.sync = sync
        self.params = sync.params
        assert [id(p) for p in self.params] == [id(p) for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.flat_p = torch.zeros_like(sync.flat)
        with torch.no_grad():
            for p, off in zip(self.params, sync.offsets):
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.step_count = 0

    def step(self, grad_scale=1.0):
        from . import ops
        self.step_count += 1
        ops._call('dx_adam_step', self.flat_p.data_ptr(), self.sync.flat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                  self.flat_p.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                  float(self.weight_decay), self.step_count, float(grad_scale), ops._st())
        ops.invalidate_packed_weights()   # raw-pointer update does not bump tensor versions

    def zero_grad(self):
        self.sync.zero_grad()
