"""Isolated GEMM timings (CUDA events, L2 flushed) at 3 / 2 / 1 tensor-core passes: which GEMMs are tensor-bound at all?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from daft_exprt_b200 import ops
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

def timed(fn, n=8):
    for _ in range(2):
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

shapes = [('conv1 fwd (planes out)', 32, 1000, 128, 1024, 3, 'planes'), ('conv2 fwd', 32, 1000, 1024, 128, 3, 'y'), ('prenet mid', 32, 1000, 1024, 1024, 3, 'y'),
          ('qkv k=1', 32, 1000, 128, 384, 1, 'y'), ('out-proj k=1', 32, 1000, 128, 128, 1, 'y')]
for name, B, S, Cin, Cout, KW, mode in shapes:
    x = torch.randn(B, S, Cin, device=dev); w = torch.randn(Cout, Cin, KW, device=dev) * 0.05; bias = torch.randn(Cout, device=dev)
    dy = torch.randn(B, S, Cout, device=dev)
    wp, wd = ops.packed(w)
    xP = ops.make_planes(x, B * S, Cin); dyP = ops.make_planes(dy, B * S, Cout)
    res = []
    for passes in (3, 2, 1):
        ops.set_gemm_passes(passes, passes)
        if mode == 'planes':
            f = lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP, emit_planes=True, want_y=False)
        else:
            f = lambda: ops.conv_gemm(x, wp, bias, B, S, x_planes=xP)
        ops.set_wgrad_deferral(True)
        g = lambda: ops.conv_wgrad(x, dy, B, S, Cin, Cout, KW, (Cout, Cin, KW), x_planes=xP, dy_planes=dyP, dbias=bias)
        tf, tw = timed(f), timed(g)
        ops._call('dx_wgrad_flush', ops._st())
        ops.set_wgrad_deferral(False)
        res.append((passes, tf, tw))
    ops.set_gemm_passes(3, 3)
    fl = 2.0 * B * S * Cin * Cout * KW
    print(f'{name:24s} ' + ' | '.join(f'p{p}: fwd {tf:6.1f} us ({fl / tf / 1e6:5.0f} TF) wgrad {tw:6.1f} us ({fl / tw / 1e6:5.0f} TF)' for p, tf, tw in res), flush=True)
