"""Per-CTA durations of the tcgen05 attention forward core (operand planes already bound) for full-length and bench-length batches."""
import os, sys
os.environ['DX_ATTN_CTA_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops, cabi
import bench
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
B, S = 32, 1000
bench_lens = bench.rank_batch(bench.CONFIGS['train'], 0)[9].to(dev)
full = torch.full((B,), S, dtype=torch.int64, device=dev)
for (H, dh, p) in [(2, 64, 0.1), (8, 16, 0.1)]:
    D = H * dh
    qkv = torch.randn(B, S, 3 * D, device=dev)
    ctx = torch.empty(B, S, D, device=dev); lse = torch.empty(B, H, S, device=dev)
    planes = ops.attention_planes(B, S, H, dh, dev)
    for name, lens in (('full', full), ('bench', bench_lens)):
        prep = lambda: ops._call('dx_attention_fwd', qkv.data_ptr(), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None,
                                 B, S, H, dh, p, 1234, ops._st())
        core = lambda: ops._call('dx_attention_fwd', None, lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), None,
                                 B, S, H, dh, p, 1234, ops._st())
        prep()
        for _ in range(3): core()
        torch.cuda.synchronize()
        evs = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); core(); e1.record(); evs.append((e0, e1))
        torch.cuda.synchronize()
        us = 1e3 * sorted(a.elapsed_time(b) for a, b in evs)[4]
        n_cta = 8 * H * B
        tr = torch.zeros(1024 + 3 * n_cta, dtype=torch.int64, device=dev)
        cabi.load().dx_debug_set_trace(tr.data_ptr())
        core(); torch.cuda.synchronize()
        cabi.load().dx_debug_set_trace(None)
        rec = tr.cpu()[1024:].view(n_cta, 3)
        live = rec[rec[:, 1] > 0]
        t0 = int(live[:, 0].min())
        dur = (live[:, 1] - live[:, 0]).double() / 1e3
        # duration by number of key tiles of the CTA's utterance
        idx = torch.nonzero(rec[:, 1] > 0).flatten()
        b_of = idx // (8 * H)
        nt = (lens.cpu()[b_of] + 127) // 128
        by = {int(k): round(float(dur[nt == k].mean()), 2) for k in nt.unique()}
        print(f'H={H} dh={dh} {name}: core {us:.1f} us; {len(live)} live CTAs, span {(int(live[:, 1].max()) - t0) / 1e3:.1f} us, mean CTA {dur.mean():.2f} us; by key tiles {by}')
        first = live[live[:, 0] < t0 + 1500]
        later = live[live[:, 0] > t0 + 20000]
        print(f'     first-wave CTAs mean {((first[:, 1] - first[:, 0]).double().mean() / 1e3):.2f} us ({len(first)}), CTAs starting after 20 us: {((later[:, 1] - later[:, 0]).double().mean() / 1e3):.2f} us ({len(later)})')
