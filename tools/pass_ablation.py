"""Pass-count ablation of the bf16x3 GEMMs: mel / loss / gradient error against the fp64 oracle and step time
for every (forward, dgrad, wgrad) pass combination of interest.  Writes gpurun_out/r2_pass_ablation.json (+ .md)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
import bench
import daft_exprt_oracle as oracle
from daft_exprt_b200 import ops, synthetic
from daft_exprt_b200.ddp import FlatAdam, FlatGradSync
from daft_exprt_b200.graph import GraphedTrainStep
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.loss import DaftExprtLoss
from daft_exprt_b200.model import DaftExprt
from helpers import l2_rel_err, scale_rel_err, targets_of

dev = torch.device('cuda', 0)
ops.set_backend('bf16x3')
N_IDS = 11
COMBOS = [(3, 3, 3), (3, 3, 0), (3, 3, 2), (3, 3, 1), (3, 2, 3), (3, 1, 3), (3, 2, 2), (3, 1, 1), (2, 2, 2), (2, 1, 1), (1, 1, 1)]   # wgrad 0 = by reduction length
if os.environ.get('DX_PA_COMBOS'):   # e.g. "3,3,3;3,3,0"
    COMBOS = [tuple(int(x) for x in c.split(',')) for c in os.environ['DX_PA_COMBOS'].split(';')]


def build(train):
    hp = default_hparams(n_speakers=N_IDS + 1)
    model = DaftExprt(hp)
    sd = synthetic.synthetic_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 1234)
    model.load_state_dict(sd)
    return model.to(dev).train(train), hp, sd


# ---- accuracy: B=6 at full length (first 6 utterances of the bench batch), eval mode, fp64 oracle on the host ----------------
full = synthetic.make_batch(32, 200, 1000, N_IDS, seed=0)
inputs = tuple(t[:6].clone() for t in full)
model, hp, sd = build(False)
crit = DaftExprtLoss(0, hp)
din = tuple(t.to(dev) for t in inputs)
ohp = oracle.OracleHParams(n_speakers=N_IDS + 1)
sd_o = {k: v.double().requires_grad_(True) for k, v in sd.items()}
in64 = tuple(t.double() if t.is_floating_point() else t for t in inputs)
ref = oracle.forward(sd_o, ohp, in64)
tot_o, _ = oracle.loss(ohp, ref, targets_of(in64), 3000)
tot_o.backward()
rows = []
for f, d, w in COMBOS:
    model.zero_grad()
    ops.set_gemm_passes(f, w)
    out = model(din)
    total = crit.forward_device(out, targets_of(din), 3000)[7]
    ops.set_gemm_passes(d, w)
    total.backward()
    ops.set_gemm_passes(3, 3)
    errs = sorted((l2_rel_err(p.grad, sd_o[n].grad), n) for n, p in model.named_parameters() if not n.startswith('gaussian_upsampling.'))
    rows.append({'passes_fwd_dgrad_wgrad': [f, d, w],
                 'mel_scale_rel': scale_rel_err(out[3][0].detach(), ref[3][0].detach()), 'mel_l2_rel': l2_rel_err(out[3][0].detach(), ref[3][0].detach()),
                 'loss_rel': abs(total.item() - tot_o.item()) / abs(tot_o.item()),
                 'grad_l2_rel_max': errs[-1][0], 'grad_l2_rel_median': errs[len(errs) // 2][0], 'grad_worst_tensor': errs[-1][1],
                 'grad_l2_rel_top5': [(round(e, 5), n) for e, n in errs[-5:]], 'grads_above_2e-3': sum(e > 2e-3 for e, _ in errs)})
    print(rows[-1], flush=True)

# ---- step time: the bench workload (B=32, train mode) as a captured graph per combination ---------------------------------------
cfg = bench.CONFIGS['train']
for row in rows:
    f, d, w = row['passes_fwd_dgrad_wgrad']
    torch.manual_seed(1234)
    model, hp, _ = build(True)
    crit = DaftExprtLoss(0, hp)
    params = list(model.parameters())
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=1e-4, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
    i_, t_, _ = model.parse_batch(0, bench.with_ids(bench.rank_batch(cfg, 0)))
    g = GraphedTrainStep(model, crit, sync, opt)
    g.before_forward = lambda f=f, w=w: ops.set_gemm_passes(f, w)
    g.before_backward = lambda d=d, w=w: ops.set_gemm_passes(d, w)
    for it in range(3):
        g.step(i_, t_, it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(8):
        g.step(i_, t_, 3 + it)
    e1.record()
    torch.cuda.synchronize()
    row['ms_per_step'] = e0.elapsed_time(e1) / 8
    ops.set_gemm_passes(3, 3)
    del g, model, opt, sync
    torch.cuda.empty_cache()
    print(row['passes_fwd_dgrad_wgrad'], 'ms/step', row['ms_per_step'], flush=True)

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump({'workload_accuracy': 'first 6 utterances of the bench batch (L<=200, T<=1000), eval mode, fp64 oracle; gradients: relative-L2 per parameter tensor '
                                '(gaussian_upsampling.* excluded: ill-conditioned sums, see tests)', 'workload_time': cfg['workload'], 'rows': rows},
          open(os.path.join(ROOT, 'gpurun_out', 'r2_pass_ablation.json'), 'w'), indent=1)
md = ['| passes fwd / dgrad / wgrad | mel scale-rel | mel l2-rel | loss rel | grad l2-rel max | grad l2-rel median | ms / step |', '|---|---|---|---|---|---|---|']
for r in rows:
    md.append(f"| {r['passes_fwd_dgrad_wgrad']} | {r['mel_scale_rel']:.1e} | {r['mel_l2_rel']:.1e} | {r['loss_rel']:.1e} | {r['grad_l2_rel_max']:.1e} "
              f"({r['grad_worst_tensor']}) | {r['grad_l2_rel_median']:.1e} | {r['ms_per_step']:.2f} |")
open(os.path.join(ROOT, 'gpurun_out', 'r2_pass_ablation.md'), 'w').write('\n'.join(md) + '\n')
print('\n'.join(md))
