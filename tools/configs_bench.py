"""BASELINE.json configs[3] and configs[4] on ONE B200 (bench.py covers configs[1]/[2]):
  [3] inference-only path (`DaftExprt.inference`, synthesize.py's call), batch = 64, variable length, eval mode
  [4] stress: training step at batch = 128, T <= 1500 (per GPU), with peak memory
Random-init weights of the reference architecture; the duration head's bias is set so that predicted durations land around
58 ms (5 frames) per phoneme, as SURVEY.md section 8(d) prescribes for the synthetic inference configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import __graft_entry__ as entry
entry.build()
from daft_exprt_b200 import ops, synthetic
from daft_exprt_b200.ddp import FlatAdam, FlatGradSync
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.loss import DaftExprtLoss
from daft_exprt_b200.model import DaftExprt

dev = torch.device('cuda', 0)
ops.set_backend('bf16x3')
N_IDS = 11
hp = default_hparams(n_speakers=N_IDS + 1)
hp.stats = {f'spk {i}': {'pitch': {'mean': 5.0 + 0.05 * i, 'std': 0.25 + 0.01 * i}} for i in range(N_IDS)}


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def config3():
    torch.manual_seed(hp.seed)
    model = DaftExprt(hp).to(dev).eval()
    with torch.no_grad():   # durations ~ 58 ms per phoneme (5 frames at hop 256 / 22050 Hz)
        lin = model.prosody_predictor.projection.linear_layer
        lin.weight.mul_(0.01)
        lin.bias.zero_()
        lin.bias[0] = 0.058
    B, L, T = 64, 200, 1000
    inputs = tuple(t.to(dev) for t in synthetic.make_inference_batch(B, L, T, N_IDS, seed=3))
    with torch.no_grad():
        ms, out = timed(lambda: model.inference(inputs, 'add', hp))
    frames = int(out[1][1].sum())
    print(json.dumps({'config': 'configs[3]: inference(), B=64, L<=200, reference mel T<=1000, pitch_transform=add, eval mode', 'ms_per_batch': ms,
                      'generated_mel_frames': frames, 'T_max_generated': int(out[1][1].max()), 'mel_frames_per_s': frames / (ms * 1e-3),
                      'audio_seconds_per_s': frames * 256 / 22050 / (ms * 1e-3)}))


def config4():
    torch.manual_seed(hp.seed)
    model = DaftExprt(hp).to(dev).train()
    crit = DaftExprtLoss(0, hp)
    params = list(model.parameters())
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=hp.initial_learning_rate, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)
    B, L, T = 128, 200, 1500
    host = synthetic.make_batch(B, L, T, N_IDS, seed=4)
    batch = tuple(host) + (['synthetic'] * B, [f'utt{i}' for i in range(B)])
    inputs, targets, _ = model.parse_batch(0, batch)
    frames = int(host[9].sum())

    def step():
        opt.zero_grad()
        out = crit.forward_device(model(inputs), targets, 1000)
        out[7].backward()
        sync.all_reduce_mean()
        opt.step()
        return out

    torch.cuda.reset_peak_memory_stats()
    ms, _ = timed(step, warmup=2, iters=5)
    print(json.dumps({'config': 'configs[4] (one GPU of it): training step, B=128, L<=200, T<=1500, dropout on, eager launches', 'ms_per_step': ms,
                      'valid_frames': frames, 'padded_frames': B * T, 'valid_mel_frames_per_s': frames / (ms * 1e-3),
                      'peak_memory_GB': torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == '__main__':
    config3()
    config4()
