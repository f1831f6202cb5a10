"""How long does the batched split-K reduction (dx_wgrad_flush, one launch per backward) take inside a training step?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from daft_exprt_b200 import ops
from daft_exprt_b200.ddp import FlatAdam, FlatGradSync
from daft_exprt_b200.graph import GraphedTrainStep
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.loss import DaftExprtLoss
from daft_exprt_b200.model import DaftExprt
ops.set_backend('bf16x3')
cfg = bench.CONFIGS['train']
hp = default_hparams(n_speakers=bench.N_SPK_IDS + 1)
torch.manual_seed(hp.seed)
model = DaftExprt(hp).cuda().train()
crit = DaftExprtLoss(0, hp)
params = list(model.parameters())
sync = FlatGradSync(params, mode='gather')
opt = FlatAdam(params, sync, lr=1e-4, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
i_, t_, _ = model.parse_batch(0, bench.with_ids(bench.rank_batch(cfg, 0)))
evs = []
orig = ops.flush_wgrad
def timed_flush():
    if len(evs) >= 6:        # the capture that follows the eager warm-up steps cannot record timing events
        return orig()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(); e1.record(); evs.append((e0, e1))
ops.flush_wgrad = timed_flush
g = GraphedTrainStep(model, crit, sync, opt, warmup=6)
g.step(i_, t_, 0)
torch.cuda.synchronize()
print('dx_wgrad_flush inside eager steps (us):', [round(a.elapsed_time(b) * 1e3, 1) for a, b in evs])
