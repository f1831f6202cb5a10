import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import __graft_entry__ as entry
entry.build()
import daft_exprt_oracle as oracle
from daft_exprt_b200 import ops, synthetic
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.loss import DaftExprtLoss
from daft_exprt_b200.model import DaftExprt
from helpers import l2_rel_err, scale_rel_err, targets_of
dev = torch.device('cuda', 0)
for nb, seed, bs in ((2, 77, (3, 33, 140)), (1, 77, (3, 33, 140)), (2, 1234, (3, 33, 140)), (2, 77, (6, 60, 300))):
    n_ids = 5
    hp = default_hparams(n_speakers=n_ids + 1)
    hp.local_prosody_predictor['nb_blocks'] = nb
    model = DaftExprt(hp)
    sd = synthetic.synthetic_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    crit = DaftExprtLoss(0, hp)
    inputs = synthetic.make_batch(*bs, n_ids, seed=12)
    din = tuple(t.to(dev) for t in inputs)
    ohp = oracle.OracleHParams(n_speakers=n_ids + 1)
    ohp.local_prosody_predictor = dict(ohp.local_prosody_predictor, nb_blocks=nb)
    sd_o = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    in64 = tuple(t.double() if t.is_floating_point() else t for t in inputs)
    ref = oracle.forward(sd_o, ohp, in64)
    tot_o, terms_o = oracle.loss(ohp, ref, targets_of(in64), 1500)
    tot_o.backward()
    res = {}
    for be in ('fp32', 'bf16x3'):
        ops.set_backend(be)
        model.zero_grad()
        out = model(din)
        total, _ = crit(out, targets_of(din), 1500)
        total.backward()
        errs = sorted(((scale_rel_err(p.grad, sd_o[n].grad), l2_rel_err(p.grad, sd_o[n].grad), n) for n, p in model.named_parameters()), reverse=True)
        print(f'nb={nb} seed={seed} bs={bs} {be}: loss {total.item():.5f} (oracle {tot_o.item():.5f}) dur pred err {scale_rel_err(out[2][0].detach(), ref[2][0].detach()):.1e}',
              'worst grads:', [(n[-45:], f'{a:.1e}', f'{b:.1e}') for a, b, n in errs[:4]], 'terms', {k: round(float(v), 4) for k, v in terms_o.items()})
