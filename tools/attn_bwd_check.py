"""dq / dk / dv errors of the attention backward vs an fp64 reference, per configuration (DX_ATTN_BWD_TC selects the kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from daft_exprt_b200 import ops
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
for (B, S, H, dh) in [(2, 150, 8, 16), (1, 257, 8, 16), (2, 400, 8, 16), (3, 70, 2, 64), (1, 300, 2, 64)]:
    D = H * dh
    g = torch.Generator().manual_seed(B * S + H)
    qkv = torch.randn(B, S, 3 * D, generator=g)
    lens = torch.randint(max(1, S // 3), S + 1, (B,), generator=g); lens[0] = S
    valid = torch.arange(S)[None, :] < lens[:, None]
    dctx = torch.randn(B, S, D, generator=g) * valid[:, :, None]
    q64 = qkv.double().clone().requires_grad_(True)
    q, k, v = q64.split(D, dim=2)
    hd = lambda t: t.reshape(B, S, H, dh).permute(0, 2, 1, 3)
    sc = (hd(q) / np.sqrt(dh)) @ hd(k).transpose(-1, -2)
    sc = sc.masked_fill(~valid[:, None, None, :], float('-inf'))
    ctx_ref = (torch.softmax(sc, -1) @ hd(v)).permute(0, 2, 1, 3).reshape(B, S, D) * valid[:, :, None]
    ctx_ref.backward(dctx.double())
    qd, ld = qkv.to(dev), lens.to(dev)
    ctx = torch.empty(B, S, D, device=dev); lse = torch.empty(B, H, S, device=dev)
    planes = ops.attention_planes(B, S, H, dh, dev)
    ops._call('dx_attention_fwd', qd.data_ptr(), ld.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None, B, S, H, dh, 0.0, 0, ops._st())
    dqkv = torch.empty(B, S, 3 * D, device=dev)
    scratch = torch.empty(ops.lib().dx_attention_bwd_scratch_bytes(B, S, H, dh), device=dev, dtype=torch.uint8)
    for rep in range(3):
        ops._call('dx_attention_bwd', qd.data_ptr(), ops._p(planes), ld.data_ptr(), ctx.data_ptr(), lse.data_ptr(), dctx.to(dev).data_ptr(),
                  dqkv.data_ptr(), scratch.data_ptr(), B, S, H, dh, 0.0, 0, ops._st())
        torch.cuda.synchronize()
        out = dqkv.cpu().double(); ref = q64.grad
        errs = [((out[..., i * D:(i + 1) * D] - ref[..., i * D:(i + 1) * D]).abs().max() / ref[..., i * D:(i + 1) * D].abs().max()).item() for i in range(3)]
        print(f'cfg {(B, S, H, dh)} lens {lens.tolist()} rep {rep}: dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e}', flush=True)
