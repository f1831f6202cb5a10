"""clock64 trace of block 0 of the tcgen05 conv-GEMM (dx_debug_set_trace): per K-step TMA issue / MMA start, per tile epilogue
start / end.  python tools/tc_trace.py [B,S,Cin,Cout,KW] [planes]   (DX_TC_DEBUG masks apply)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops, cabi
ops.set_backend('bf16x3')
ops.set_gemm_passes(int(os.environ.get('DX_PASSES', '3')), 3)
dev = torch.device('cuda', 0)
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '32,1000,128,1024,3').split(','))
planes = len(sys.argv) > 2
B, S, Cin, Cout, KW = shape
x = torch.randn(B, S, Cin, device=dev); w = torch.randn(Cout, Cin, KW, device=dev) * 0.05; bias = torch.randn(Cout, device=dev)
wp, _ = ops.packed(w)
xP = ops.make_planes(x, B * S, Cin)
run = (lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP, emit_planes=True, want_y=False)) if planes else \
      (lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP))
for _ in range(2): run()
torch.cuda.synchronize()
tr = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
cabi.load().dx_debug_set_trace(tr.data_ptr())
run()
torch.cuda.synchronize()
cabi.load().dx_debug_set_trace(None)
t = tr.cpu().view(4, 256)
t0 = int(t[t > 0].min())
names = ['tma_issued', 'mma_full', 'epi_tfull', 'epi_done']
rows = [[int(a) - t0 for a in t[r] if a > 0] for r in range(4)]
tag = f'{shape} planes={planes} mask={os.environ.get("DX_TC_DEBUG", "0")}'
for r in range(4):
    print(tag, names[r], rows[r][:28])
n = min(len(rows[2]), len(rows[3]))
print(tag, 'epilogue duration per tile', [rows[3][i] - rows[2][i] for i in range(n)][:16])
print(tag, 'tile period (epi start)  ', [rows[2][i + 1] - rows[2][i] for i in range(n - 1)][:16])
print(tag, 'k-step period (mma)      ', [rows[1][i + 1] - rows[1][i] for i in range(min(len(rows[1]) - 1, 30))])
