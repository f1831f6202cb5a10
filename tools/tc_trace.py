import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops, cabi
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '32,1000,128,1024,3').split(','))
B, S, Cin, Cout, KW = shape
x = torch.randn(B, S, Cin, device=dev); w = torch.randn(Cout, Cin, KW, device=dev) * 0.05; bias = torch.randn(Cout, device=dev)
wp, _ = ops.packed(w)
for _ in range(2): ops.conv_gemm(x, wp, bias, B, S, relu=True)
torch.cuda.synchronize()
tr = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
cabi.load().dx_debug_set_trace(tr.data_ptr())
ops.conv_gemm(x, wp, bias, B, S, relu=True)
torch.cuda.synchronize()
cabi.load().dx_debug_set_trace(None)
t = tr.cpu().view(4, 256)
t0 = int(t[t > 0].min())
names = ['tma_issued', 'mma_full', 'epi_tfull', 'epi_done']
for r in range(4):
    v = [int(a) - t0 for a in t[r] if a > 0][:40]
    print(shape, names[r], v)
