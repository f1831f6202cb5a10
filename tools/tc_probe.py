"""Bring-up probe for the tcgen05 GEMM: one small conv-GEMM through the C-ABI, compared with torch (run per DX_TC_DEBUG mask)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops
backend = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
ops.set_backend(backend)
dev = torch.device('cuda', 0)
B, S, Cin, Cout, KW = 2, 200, 128, 256, 3
g = torch.Generator().manual_seed(0)
x = torch.randn(B, S, Cin, generator=g).to(dev)
w = (torch.randn(Cout, Cin, KW, generator=g) / 20).to(dev)
b = torch.randn(Cout, generator=g).to(dev)
wp, wd = ops.packed(w)
y = ops.conv_gemm(x, wp, b, B, S)
torch.cuda.synchronize()
ref = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=1).transpose(1, 2)
err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
print(f'probe[{backend}] DX_TC_DEBUG={os.environ.get("DX_TC_DEBUG", "0")}: sync ok, scale-rel err {err:.3e}')
dy = torch.randn(B, S, Cout, generator=g).to(dev)
dw, db = ops.conv_wgrad(x, dy, B, S, Cin, Cout, KW, (Cout, Cin, KW))
torch.cuda.synchronize()
wr = w.double().clone().requires_grad_(True)
torch.nn.functional.conv1d(x.double().transpose(1, 2), wr, None, padding=1).transpose(1, 2).backward(dy.double())
print(f'probe[{backend}] wgrad err {(dw.double() - wr.grad).abs().max().item() / wr.grad.abs().max().item():.3e}')
