"""N-rank check of ddp.FusedShardedAdam against NCCL all-reduce + FlatAdam (run under torchrun, N >= 2):
identical random 'gradients' per rank -> both paths must give the same parameters / moments after three optimiser steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import __graft_entry__ as entry
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
if rank == 0:
    entry.build()
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
dist.barrier()
from daft_exprt_b200.ddp import FlatAdam, FlatGradSync, FusedShardedAdam

def make():
    torch.manual_seed(5)
    params = [torch.nn.Parameter(torch.randn(*s, device=dev)) for s in ((1024, 128, 3), (384, 128), (1000,), (7,), (128, 1024, 3), (3, 5))]
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
    return params, sync, opt

pa, sa, oa = make()          # NCCL all-reduce + Adam
pb, sb, ob = make()          # fused
fused = FusedShardedAdam(sb, ob)
if rank == 0:
    print('fused exchange mode:', fused.mode, '| shard', fused.begin, fused.n, 'of', sb.padded_numel, flush=True)
for step in range(3):
    g = torch.Generator(device=dev).manual_seed(100 * step + rank)
    grads = [torch.randn(p.shape, device=dev, generator=g) for p in pa]
    for p, q, gr in zip(pa, pb, grads):
        p.grad = gr.clone()
        q.grad = gr.clone()
    sa.all_reduce_mean()
    oa.step()
    fused.step()
    torch.cuda.synchronize()
m_full, v_full = fused.full_moments()
err_p = max(float((p.data - q.data).abs().max() / p.data.abs().max()) for p, q in zip(pa, pb))
err_m = float((oa.m - m_full).abs().max() / oa.m.abs().max())
err_v = float((oa.v - v_full).abs().max() / oa.v.abs().max())
same = torch.tensor([float(ob.flat_p.double().sum())], device=dev, dtype=torch.float64)
lo, hi = same.clone(), same.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
print(f'rank {rank}: params rel err {err_p:.2e}, exp_avg {err_m:.2e}, exp_avg_sq {err_v:.2e}; parameter checksum spread over ranks {float(hi - lo):.3e}', flush=True)
assert err_p < 1e-5 and err_m < 1e-5 and err_v < 1e-5 and float(hi - lo) == 0.0
# timing: 58.9 MB bucket like the real model
n = 14727153
big = [torch.nn.Parameter(torch.randn(n, device=dev))]
s1 = FlatGradSync(big, mode='gather'); o1 = FlatAdam(big, s1)
big2 = [torch.nn.Parameter(torch.randn(n, device=dev))]
s2 = FlatGradSync(big2, mode='gather'); o2 = FlatAdam(big2, s2)
f2 = FusedShardedAdam(s2, o2)
big[0].grad = torch.randn(n, device=dev); big2[0].grad = big[0].grad.clone()
def timed(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
t_nccl = timed(lambda: (s1.all_reduce_mean(), o1.step()))
t_fused = timed(lambda: f2.step())
if rank == 0:
    print(f'58.9 MB bucket, {world} ranks: NCCL all-reduce + Adam {t_nccl:.1f} us | fused kernel (+ 2 barriers) {t_fused:.1f} us', flush=True)
dist.destroy_process_group()
