"""Per-CTA start/end (globaltimer) and SM id of the tcgen05 attention forward at the bench batch (DX_ATTN_CTA_TRACE=1)."""
import os, sys
os.environ['DX_ATTN_CTA_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops, cabi
import bench
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
B, S = 32, 1000
lens = bench.rank_batch(bench.CONFIGS['train'], 0)[9].to(dev)
for (H, dh, p) in [(2, 64, 0.1), (8, 16, 0.1)]:
    D = H * dh
    qkv = torch.randn(B, S, 3 * D, device=dev)
    ctx = torch.empty(B, S, D, device=dev); lse = torch.empty(B, H, S, device=dev)
    planes = ops.attention_planes(B, S, H, dh, dev)
    run = lambda: ops._call('dx_attention_fwd', qkv.data_ptr(), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None,
                            B, S, H, dh, p, 1234, ops._st())
    for _ in range(2): run()
    torch.cuda.synchronize()
    n_cta = 8 * H * B
    tr = torch.zeros(1024 + 3 * n_cta, dtype=torch.int64, device=dev)
    cabi.load().dx_debug_set_trace(tr.data_ptr())
    run(); torch.cuda.synchronize()
    cabi.load().dx_debug_set_trace(None)
    rec = tr.cpu()[1024:].view(n_cta, 3)
    live = rec[rec[:, 1] > 0]
    t0 = int(rec[:, 0][rec[:, 0] > 0].min())
    dur = (live[:, 1] - live[:, 0]).double()
    print(f'H={H} dh={dh}: {n_cta} CTAs, {len(live)} live; kernel span {(int(live[:, 1].max()) - t0) / 1e3:.1f} us; CTA duration mean {dur.mean() / 1e3:.2f} us, '
          f'max {dur.max() / 1e3:.2f} us, sum/148 = {dur.sum() / 148 / 1e3:.1f} us; first live CTA start +{(int(live[:, 0].min()) - t0) / 1e3:.2f} us')
    for sm in (0, 1, 77):
        mine = live[live[:, 2] == sm]
        mine = mine[mine[:, 0].argsort()]
        print(f'   SM {sm}:', [(round((int(a) - t0) / 1e3, 1), round((int(b) - t0) / 1e3, 1)) for a, b, _ in mine.tolist()])
