"""Is the training step host-bound?  Times (a) the CPU enqueue of one step, (b) the device time of the step, and (c) a CUDA-graph
replay of the same step (naive capture: seeds / Adam step count baked in -- timing probe only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import bench
from daft_exprt_b200 import ops
from daft_exprt_b200.ddp import FlatAdam, FlatGradSync
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.loss import DaftExprtLoss
from daft_exprt_b200.model import DaftExprt

dev = torch.device('cuda', 0)
ops.set_backend('bf16x3')
hp = default_hparams(n_speakers=bench.N_SPK_IDS + 1)
torch.manual_seed(hp.seed)
model = DaftExprt(hp).to(dev).train()
crit = DaftExprtLoss(0, hp)
params = list(model.parameters())
sync = FlatGradSync(params, mode='gather')
opt = FlatAdam(params, sync, lr=hp.initial_learning_rate, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
inputs, targets, _ = model.parse_batch(0, bench.with_ids(bench.rank_batch(bench.CONFIGS['train'], 0)))

def step(it):
    opt.zero_grad()
    out = crit.forward_device(model(inputs), targets, it)
    out[7].backward()
    sync.all_reduce_mean()
    opt.step()
    return out

for i in range(3):
    step(i)
torch.cuda.synchronize()
N = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(N):
    step(3 + i)
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f'eager: enqueue {t_enq / N * 1e3:.2f} ms/step, device {e0.elapsed_time(e1) / N:.2f} ms/step, wall {t_all / N * 1e3:.2f}')

try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(2):
            step(20 + i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    opt.zero_grad()
    with torch.cuda.graph(g):
        out = step(30)
    torch.cuda.synchronize()
    for i in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for i in range(N):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f'graph replay: device {e0.elapsed_time(e1) / N:.2f} ms/step; loss terms {out.tolist()}')
except Exception as ex:
    import traceback; traceback.print_exc()
