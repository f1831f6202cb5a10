"""How far do two runs of the SAME training trajectory drift apart (fp32 atomics ordering -> Adam)?  eager vs eager, graph vs graph, eager vs graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
from daft_exprt_b200 import ops, synthetic
from daft_exprt_b200.ddp import FlatAdam, FlatGradSync
from daft_exprt_b200.graph import GraphedTrainStep
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.loss import DaftExprtLoss
from daft_exprt_b200.model import DaftExprt
dev = torch.device('cuda', 0)
ops.set_backend('bf16x3')
batch = synthetic.make_batch(4, 40, 260, 5, seed=3) + ([], [])
lr_of = lambda it: 1e-4 * (1 + it % 3)
iters = [2000, 2001, 2002, 2003, 2004, 2005]

def objs():
    hp = default_hparams(n_speakers=6)
    model = DaftExprt(hp)
    sd = synthetic.synthetic_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 1234)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    crit = DaftExprtLoss(0, hp)
    params = list(model.parameters())
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=1e-3, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
    return model, crit, sync, opt

def run(kind):
    model, crit, sync, opt = objs()
    inputs, targets, _ = model.parse_batch(0, batch)
    out = []
    if kind == 'graph':
        g = GraphedTrainStep(model, crit, sync, opt, lr_schedule=lr_of)
        for it in iters:
            out.append(g.step(inputs, targets, it).cpu().numpy().copy())
    else:
        for it in iters:
            opt.lr = lr_of(it)
            opt.zero_grad()
            o = crit.forward_device(model(inputs), targets, it)
            o[7].backward()
            sync.all_reduce_mean()
            opt.step()
            out.append(o.detach().cpu().numpy().copy())
    return np.stack(out), opt.flat_p.detach().cpu().numpy().copy()

res = {k: run(k.split('_')[0]) for k in ('eager_a', 'eager_b', 'graph_a', 'graph_b')}
def cmp(a, b):
    la, lb = res[a][0], res[b][0]
    rel = np.abs(la - lb) / np.maximum(np.abs(la), 1e-9)
    return [f'{r.max():.1e}' for r in rel], f'param l2 {np.linalg.norm(res[a][1] - res[b][1]) / np.linalg.norm(res[a][1]):.2e}'
for a, b in (('eager_a', 'eager_b'), ('graph_a', 'graph_b'), ('eager_a', 'graph_a')):
    print(a, 'vs', b, 'max rel loss-term diff per step:', *cmp(a, b))
