"""clock64 trace of CTA (0,0,0) of the tcgen05 attention backward (DX_ATTN_BWD_TC=1): per query tile, softmax warp 2 lane 0."""
import os, sys
os.environ['DX_ATTN_BWD_TC'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops, cabi
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
B, S = 32, 1000
lens = torch.full((B,), S, dtype=torch.int64, device=dev)
for (H, dh, p) in [(2, 64, 0.1), (8, 16, 0.1)]:
    D = H * dh
    qkv = torch.randn(B, S, 3 * D, device=dev)
    ctx = torch.empty(B, S, D, device=dev); lse = torch.empty(B, H, S, device=dev)
    planes = ops.attention_planes(B, S, H, dh, dev)
    ops._call('dx_attention_fwd', qkv.data_ptr(), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None, B, S, H, dh, p, 1234, ops._st())
    dctx = torch.randn(B, S, D, device=dev)
    dqkv = torch.empty(B, S, 3 * D, device=dev)
    scratch = torch.empty(ops.lib().dx_attention_bwd_scratch_bytes(B, S, H, dh), device=dev, dtype=torch.uint8)
    run = lambda: ops._call('dx_attention_bwd', qkv.data_ptr(), ops._p(planes), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), dctx.data_ptr(),
                            dqkv.data_ptr(), scratch.data_ptr(), B, S, H, dh, p, 1234, ops._st())
    for _ in range(2): run()
    torch.cuda.synchronize()
    tr = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
    cabi.load().dx_debug_set_trace(tr.data_ptr())
    run(); torch.cuda.synchronize()
    cabi.load().dx_debug_set_trace(None)
    t = tr.cpu().view(4, 256)
    t0 = int(t[t > 0].min())
    rows = [[int(a) - t0 for a in t[r] if a > 0] for r in range(4)]
    n = len(rows[0])
    if dh == 16:   # pipelined kernel: per half-step {begin, scores available, arithmetic done, published}
        print(f'H={H} dh={dh}: {n} half-steps (pipelined kernel)')
        print('  half-step period     ', [rows[0][i + 1] - rows[0][i] for i in range(n - 1)])
        print('  bar.sync + wait S^T  ', [rows[1][i] - rows[0][i] for i in range(n)])
        print('  ld + arithmetic + st ', [rows[2][i] - rows[1][i] for i in range(n)])
        print('  guard + staging + pub', [rows[3][i] - rows[2][i] for i in range(n)])
        print('  drain (after publish)', [rows[0][i + 1] - rows[3][i] for i in range(n - 1)])
        continue
    print(f'H={H} dh={dh}: {n} query tiles')
    print('  s_full seen          ', rows[0])
    print('  wait for S^T/dP^T    ', [rows[0][i] - rows[3][i - 1] for i in range(1, n)])
    print('  softmax (ld..publish)', [rows[1][i] - rows[0][i] for i in range(n)])
    print('  wait for dQ          ', [rows[2][i] - rows[1][i] for i in range(n)])
    print('  drain dQ             ', [rows[3][i] - rows[2][i] for i in range(n)])
