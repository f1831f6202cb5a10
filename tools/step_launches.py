"""One eager training step (bench.py workload) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/step_launches.py"""
import os, sys
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
ops.set_backend(os.environ.get('DX_BACKEND', 'bf16x3'))
hp = default_hparams(n_speakers=bench.N_SPK_IDS + 1)
torch.manual_seed(hp.seed)
model = DaftExprt(hp).to(dev).train()
crit = DaftExprtLoss(0, hp)
params = list(model.parameters())
sync = FlatGradSync(params, mode='gather')
opt = FlatAdam(params, sync, lr=hp.initial_learning_rate, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
cfg = bench.CONFIGS[os.environ.get('DX_CONFIG', 'train')]
inputs, targets, _ = model.parse_batch(0, bench.with_ids(bench.rank_batch(cfg, 0)))


def step(it):
    opt.zero_grad()
    out = crit.forward_device(model(inputs), targets, it)
    out[7].backward()
    sync.all_reduce_mean()
    opt.step()


for i in range(2):
    step(i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
