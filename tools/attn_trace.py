"""clock64 trace of CTA (0,0,0) of the tcgen05 attention forward: per step, softmax warp 2: s_full seen / scores loaded / step done;
MMA thread: S(t) issued."""
import os, sys
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
    run = lambda: ops._call('dx_attention_fwd', qkv.data_ptr(), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None,
                            B, S, H, dh, p, 1234, ops._st())
    for _ in range(2): run()
    torch.cuda.synchronize()
    tr = torch.zeros(4 * 256, dtype=torch.int64, device=dev)
    cabi.load().dx_debug_set_trace(tr.data_ptr())
    run(); torch.cuda.synchronize()
    cabi.load().dx_debug_set_trace(None)
    t = tr.cpu().view(4, 256)
    cta = [int(t[3][i]) for i in (200, 201, 202, 204, 205)]
    t[3][200:] = 0
    t0 = int(t[t > 0].min())
    print(f'  CTA entry -> TMEM ready {cta[1] - cta[0]} cyc; entry -> first MMA {t0 - cta[0]} cyc; entry -> exit {cta[2] - cta[0]} cyc = {cta[4] - cta[3]} ns  ({(cta[2] - cta[0]) / max(cta[4] - cta[3], 1):.2f} GHz)')
    rows = [[int(a) - t0 for a in t[r] if a > 0] for r in range(4)]
    print(f'H={H} dh={dh}: 16 steps (8 pass A + 8 pass B)')
    for name, r in zip(['s_full seen', 'scores loaded', 'step done', 'MMA issued S(t)'], rows):
        print(f'  {name:16s}', r[:16])
    print('  step period     ', [rows[0][i + 1] - rows[0][i] for i in range(min(15, len(rows[0]) - 1))])
    print('  ld latency      ', [rows[1][i] - rows[0][i] for i in range(min(16, len(rows[0])))])
    print('  compute+st      ', [rows[2][i] - rows[1][i] for i in range(min(16, len(rows[0])))])
