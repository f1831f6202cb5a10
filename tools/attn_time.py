"""Time the attention forward (and backward) through the C-ABI at the bench shapes.  DX_ATTN_TC=0 -> mma.sync forward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from daft_exprt_b200 import ops
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
import bench
lens = bench.rank_batch(bench.CONFIGS['train'], 0)[9].to(dev)   # output_lengths of the bench batch
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for (B, S, H, dh, p) in [(32, 1000, 2, 64, 0.1), (32, 1000, 8, 16, 0.1), (32, 1000, 2, 64, 0.0), (32, 1000, 8, 16, 0.0)]:
    D = H * dh
    qkv = torch.randn(B, S, 3 * D, device=dev)
    ctx = torch.empty(B, S, D, device=dev); lse = torch.empty(B, H, S, device=dev)
    planes = ops.attention_planes(B, S, H, dh, dev)
    ctxP = torch.empty(2, B * S, D, device=dev, dtype=torch.bfloat16)
    run = lambda: ops._call('dx_attention_fwd', qkv.data_ptr(), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes), ops._p(ctxP),
                            B, S, H, dh, p, 1234, ops._st())
    for _ in range(3): run()
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    us = 1e3 * sum(ms[:5]) / 5
    fl = 4.0 * float((lens.double() ** 2).sum()) * D   # valid-only algorithmic flops
    dctx = torch.randn(B, S, D, device=dev)
    dqkv = torch.empty(B, S, 3 * D, device=dev)
    scratch = torch.empty(ops.lib().dx_attention_bwd_scratch_bytes(B, S, H, dh), device=dev, dtype=torch.uint8)
    runb = lambda: ops._call('dx_attention_bwd', qkv.data_ptr(), ops._p(planes), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), dctx.data_ptr(),
                             dqkv.data_ptr(), scratch.data_ptr(), B, S, H, dh, p, 1234, ops._st())
    for _ in range(3): runb()
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); runb(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    msb = sorted(a.elapsed_time(b) for a, b in evs)
    usb = 1e3 * sum(msb[:5]) / 5
    print(f'attention bwd (delta + memset + prep + core) B={B} S={S} H={H} dh={dh} p={p}: {usb:7.1f} us  {2.5 * fl / usb / 1e6:6.1f} TFLOP/s alg  bwd_tc={os.environ.get("DX_ATTN_BWD_TC", "0")}')
    print(f'attention fwd (prep + core) B={B} S={S} H={H} dh={dh} p={p}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s alg (valid keys/queries)  tc={os.environ.get("DX_ATTN_TC", "1")}')
