"""A/B timing of the captured training step under feature switches, all variants on the SAME box, interleaved twice.
   python tools/ab_step.py            (parent)     |   python tools/ab_step.py child   (one variant, switches from the environment)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    'all_on': {},
    'no_fused_ln': {'DX_AB_FUSED_LN': '0'},
    'no_fused_inproj': {'DX_AB_FUSED_INPROJ': '0'},
    'no_pipe64': {'DX_ATTN_BWD_PIPE64': '0'},
    'all_off': {'DX_AB_FUSED_LN': '0', 'DX_AB_FUSED_INPROJ': '0', 'DX_ATTN_BWD_PIPE64': '0'},
    'two_pass_softmax': {'DX_ATTN_FWD_ONLINE': '0'},
    'fwd64_2cta': {'DX_ATTN_FWD64_2CTA': '1'},
    'no_wgrad_three_tap': {'DX_WGRAD_HALO3': '0'},
}
if os.environ.get('DX_AB_ONLY'):   # comma list of variant names
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in os.environ['DX_AB_ONLY'].split(',')}
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    import bench
    from daft_exprt_b200 import ops
    from daft_exprt_b200.ddp import FlatAdam, FlatGradSync
    from daft_exprt_b200.graph import GraphedTrainStep
    from daft_exprt_b200.hparams import default_hparams
    from daft_exprt_b200.loss import DaftExprtLoss
    from daft_exprt_b200.model import DaftExprt
    ops.set_backend('bf16x3')
    ops.fused_ln(os.environ.get('DX_AB_FUSED_LN', '1') != '0')
    ops.fused_inproj(os.environ.get('DX_AB_FUSED_INPROJ', '1') != '0')
    cfg = bench.CONFIGS['train']
    hp = default_hparams(n_speakers=bench.N_SPK_IDS + 1)
    torch.manual_seed(hp.seed)
    model = DaftExprt(hp).cuda().train()
    crit = DaftExprtLoss(0, hp)
    params = list(model.parameters())
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=1e-4, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
    i_, t_, _ = model.parse_batch(0, bench.with_ids(bench.rank_batch(cfg, 0)))
    g = GraphedTrainStep(model, crit, sync, opt)
    for it in range(4):
        g.step(i_, t_, it)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(10):
            g.step(i_, t_, 4 + it)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(json.dumps({'ms_per_step': best, 'launches': g.cache[next(iter(g.cache))].launches}))
else:
    import __graft_entry__ as entry
    entry.build()
    res = {k: [] for k in VARIANTS}
    for rnd in range(2):
        for name, env in VARIANTS.items():
            out = subprocess.run([sys.executable, __file__, 'child'], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
            try:
                res[name].append(json.loads(out.stdout.strip().splitlines()[-1]))
            except Exception:
                res[name].append({'error': out.stderr[-300:]})
    for name, r in res.items():
        print(name, r)
