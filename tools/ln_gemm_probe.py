"""Isolated launches of the GEMM + LayerNorm epilogue kernels at the bench shapes (for ncu / timing):
   DX_PROBE=conv2|outproj  python tools/ln_gemm_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import bench
from daft_exprt_b200 import ops
dev = torch.device('cuda', 0)
ops.set_backend('bf16x3')
cfg = bench.CONFIGS['train']
B, S, D = cfg['B'], cfg['T'], 128
which = os.environ.get('DX_PROBE', 'conv2')
Cin, KW = (1024, 3) if which == 'conv2' else (128, 1)
lens = bench.rank_batch(cfg, 0)[9].to(dev)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, S, Cin, generator=g).to(dev)
w = (torch.randn(D, Cin, KW, generator=g) * 0.03).to(dev)
bias = torch.randn(D, generator=g).to(dev)
res = torch.randn(B, S, D, generator=g).to(dev)
lw, lb = torch.ones(D, device=dev), torch.zeros(D, device=dev)
film = torch.randn(B, 2 * D, generator=g).to(dev)
wp, _ = ops.packed(w)
xP = ops.make_planes(x, B * S, Cin)
p = float(os.environ.get('DX_P', '0.1'))
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2] * 1e3

fused = lambda: ops.gemm_ln(xP, wp, bias, res, lw, lb, film if which == 'conv2' else None, 2 * D, lens, B, S, p_in=p, seed_in=77)
def split():
    o = ops.conv_gemm(None if False else x, wp, bias, B, S, x_planes=xP, lens=lens)
    return ops.ln_fwd(o, res, lw, lb, film if which == 'conv2' else None, 2 * D, lens, B, S, D, p_in=p, seed_in=77, emit_planes=True)
gemm_only = lambda: ops.conv_gemm(x, wp, bias, B, S, x_planes=xP, lens=lens)
print(which, 'fused %.1f us' % timed(fused), '| gemm + ln_fwd %.1f us' % timed(split), '| gemm only %.1f us' % timed(gemm_only))
