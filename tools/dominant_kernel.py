"""The dominant kernel alone, exactly as bench.py times it (for `ncu --set full -k regex:gemm_tc_kernel --launch-skip 3 -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import bench
from daft_exprt_b200 import ops
ops.set_backend('bf16x3')
dev = torch.device('cuda', 0)
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
flops, t = bench.time_dominant_kernel(bench.CONFIGS['train'], dev, flush)
print(f'{flops / t / 1e12:.1f} algorithmic TFLOP/s, {t * 1e6:.1f} us per launch')
