"""Bisect what bounds the tcgen05 conv-GEMM: time a few shapes under the DX_TC_DEBUG masks (1 no MMA, 2 no TMA loads,
4 no TMEM loads, 8 no TMA stores), each mask in its own process (the mask is read once).  Results are wrong by design
under a non-zero mask; only the timings matter."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    from daft_exprt_b200 import ops
    ops.set_backend('bf16x3')
    dev = torch.device('cuda', 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def timeit(run, iters=10):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        evs = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        return 1e3 * sum(ms[:5]) / 5

    out = []
    for (B, S, Cin, Cout, KW, planes) in [(32, 1000, 128, 1024, 3, True), (32, 1000, 128, 1024, 3, False), (32, 1000, 1024, 128, 3, False),
                                          (32, 1000, 128, 384, 1, False), (32, 1000, 1024, 1024, 3, False)]:
        x = torch.randn(B, S, Cin, device=dev)
        w = torch.randn(Cout, Cin, KW, device=dev) * 0.05
        bias = torch.randn(Cout, device=dev)
        wp, _ = ops.packed(w)
        xP = ops.make_planes(x, B * S, Cin)
        if planes:
            run = lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP, emit_planes=True, want_y=False)
        else:
            run = lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP)
        us = timeit(run)
        out.append(f'{Cin}->{Cout}k{KW}{"P" if planes else "F"}:{us:7.1f}us')
    print(f'mask={os.environ.get("DX_TC_DEBUG", "0"):>2}  ' + '  '.join(out), flush=True)
else:
    for mask in (sys.argv[1:] or ['0', '1', '2', '4', '8', '12', '3', '15']):
        env = dict(os.environ, DX_TC_DEBUG=mask)
        subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=env, timeout=120)
