"""Bring-up: where do the fused in-projection planes differ from the conversion-pass planes?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
from daft_exprt_b200 import ops
dev = torch.device('cuda', 0)
ops.set_backend('bf16x3')
for (B, S, H, dh) in [(2, 300, 2, 64), (3, 129, 8, 16), (2, 70, 4, 32)]:
    D = H * dh
    g = torch.Generator().manual_seed(S * H)
    x = torch.randn(B, S, D, generator=g).to(dev)
    w = (torch.randn(3 * D, D, generator=g) / np.sqrt(D)).to(dev)
    b = torch.randn(3 * D, generator=g).to(dev)
    lens = torch.randint(1, S + 1, (B,), generator=g); lens[0] = S; lens = lens.to(dev)
    x = x * (torch.arange(S, device=dev)[None, :] < lens[:, None])[:, :, None]
    wp, _ = ops.packed(w)
    xP = ops.make_planes(x, B * S, D)
    Sp = (S + 63) // 64 * 64
    res = []
    for fused in (False, True):
        planes = ops.attention_planes(B, S, H, dh, dev); planes.fill_(0x7f)
        ctx = torch.empty(B, S, D, device=dev); lse = torch.empty(B, H, S, device=dev)
        if fused:
            ops._call('dx_inproj_head_planes', ops._p(xP), ops._p(wp.planes), ops._p(b), ops._p(planes), ops._p(lens), B, S, D, H, dh, ops._st())
            qkv = None
        else:
            qkv = ops.conv_gemm(x, wp, b, B, S, x_planes=xP, lens=lens)
        ops._call('dx_attention_fwd', ops._p(qkv), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None, B, S, H, dh, 0.0, 0, ops._st())
        torch.cuda.synchronize()
        n = 2 * B * 3 * H * Sp * dh
        off = (-planes.data_ptr()) % 256
        R = planes[off:off + 2 * n].view(torch.bfloat16).view(2, B, 3 * H, Sp, dh).float().cpu()
        res.append((R, ctx.cpu(), lse.cpu()))
    d = (res[0][0] != res[1][0])
    print(f'cfg B={B} S={S} H={H} dh={dh} lens={lens.tolist()} Sp={Sp}: mismatching elements {int(d.sum())} of {d.numel()}')
    if d.any():
        idx = d.nonzero()
        for k in range(5):
            print('   per-dim unique values of mismatch index dim', k, ':', idx[:, k].unique().tolist()[:40])
        i = tuple(idx[0].tolist())
        print('   first mismatch', i, 'unfused', res[0][0][i].item(), 'fused', res[1][0][i].item())
        pl, bb, slot, ss, dd = i
        print('   unfused row:', res[0][0][pl, bb, slot, ss, :8].tolist())
        print('   fused   row:', res[1][0][pl, bb, slot, ss, :8].tolist())
    print('   ctx equal', torch.equal(res[0][1], res[1][1]), 'max diff', (res[0][1] - res[1][1]).abs().max().item())
