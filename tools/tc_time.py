"""Time one conv-GEMM shape through the C-ABI (back-to-back launches, CUDA events around the loop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops
backend = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
shapes = [(32, 1000, 128, 1024, 3), (32, 1000, 1024, 128, 3), (32, 1000, 1024, 1024, 3), (1, 32000, 128, 384, 1)]
if len(sys.argv) > 2:
    shapes = [tuple(int(v) for v in sys.argv[2].split(','))]
ops.set_backend(backend)
dev = torch.device('cuda', 0)
for (B, S, Cin, Cout, KW) in shapes:
    x = torch.randn(B, S, Cin, device=dev)
    w = torch.randn(Cout, Cin, KW, device=dev) * 0.05
    bias = torch.randn(Cout, device=dev)
    wp, wd = ops.packed(w)
    for _ in range(3):
        y = ops.conv_gemm(x, wp, bias, B, S, relu=True)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = ops.conv_gemm(x, wp, bias, B, S, relu=True)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    fl = 2.0 * B * S * Cin * Cout * KW
    print(f'[{backend} dbg={os.environ.get("DX_TC_DEBUG","0")}] B={B} S={S} Cin={Cin} Cout={Cout} KW={KW}: {us:8.1f} us/launch (incl. split)  {fl/us/1e6:7.1f} TFLOP/s alg')
    dy = torch.randn(B, S, Cout, device=dev)
    for _ in range(2):
        ops.conv_wgrad(x, dy, B, S, Cin, Cout, KW, (Cout, Cin, KW))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        ops.conv_wgrad(x, dy, B, S, Cin, Cout, KW, (Cout, Cin, KW))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f'    wgrad: {us:8.1f} us/launch (incl. splits+reduce+colsum)  {fl/us/1e6:7.1f} TFLOP/s alg')
