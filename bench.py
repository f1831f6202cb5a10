#!/usr/bin/env python
"""Headline benchmark: valid mel-frames/s of the Daft-Exprt training step (forward + loss + backward + gradient all-reduce +
fused Adam) on synthetic batches of BASELINE.json configs[1]/[2] (11-speaker hparams, B=32 per GPU, L<=200, T<=1000, 80 mels).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's sm_100a path (configs[1]/[2])
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps 1 --warmup 0          # the reference algorithm on the host CPU (oracle port)
    python bench.py --config infer64                               # configs[3]: inference(), B=64, variable length
    python bench.py --config stress [--gpus 8 under torchrun]      # configs[4]: B=128 per GPU, T<=1500, HBM report

Prints ONE JSON line (rank 0):
  value            whole-job valid frames/s, inputs resident in HBM, fixed batch shape (the config's maxima), CUDA-graph replay
  e2e              the same metric through the public API with HOST buffers: every step takes a DIFFERENT pinned host batch (6 batches
                   of 4 distinct bucketed (L_max, T_max) shapes, data.BucketedCollate -> FlatBatch), ONE H2D copy per step via
                   parse_batch on a side stream, graph replay keyed on the bucket shape, loss read-back every step;
                   `graph_hit_rate`, the fixed-shape e2e and the eager (no graph) step time are reported beside it
  roofline         the dominant kernel (tcgen05 conv-GEMM) timed live with CUDA events; `roofline_attention`: the attention kernels
                   (north_star: throughput as a fraction of the attention-GEMM roofline)
  cpu_baseline     the oracle port of the reference on this box's host cores (bounded sample, all cores, every N)
  gpu_eager_baseline  the same oracle port (plain PyTorch eager ops, fp32, TF32 off) on cuda:0: the stock-PyTorch-on-B200 bar
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

UNIT = 'valid mel-frames/s'
N_SPK_IDS = 11
CONFIGS = {
    'train': dict(B=32, L=200, T=1000, kind='train', metric='mel_frames_per_sec_train_step',
                  workload='configs[1]/[2]: 11-speaker LJ+ESD hparams, B=32 per GPU, L<=200 phonemes, T<=1000 frames, 80 mels; step = '
                           'forward + DaftExprtLoss + backward + flat-bucket grad all-reduce (N>1) + fused Adam; train mode, dropout 0.1'),
    'stress': dict(B=128, L=200, T=1500, kind='train', metric='mel_frames_per_sec_train_step',
                   workload='configs[4]: stress, B=128 per GPU, L<=200 phonemes, T<=1500 frames, 80 mels; same training step; train mode'),
    'infer64': dict(B=64, L=200, T=1000, kind='infer', metric='mel_frames_per_sec_inference',
                    workload='configs[3]: DaftExprt.inference() (synthesize.py path), B=64 variable length, L<=200 phonemes, reference '
                             'mel T<=1000, pitch_transform=add, eval mode; frames = GENERATED mel frames'),
}
# e2e: (L_max, T_max) of the host batches cycled through the public API, as fractions of the config maxima
E2E_SHAPES = ((1.0, 1.0), (0.96, 0.97), (0.94, 0.93), (1.0, 0.89), (0.95, 0.88), (0.98, 0.99))


def host_threads():
    """Cores this process may use (affinity-aware).  torchrun exports OMP_NUM_THREADS=1: the CPU legs set the thread count
    explicitly so that the reference arm is never starved (round-1 N>1 ratios were void for that reason)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max([int(s[1]) for s in self.samples if s[1].isdigit()] or [0])
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': reasons, 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# synthetic batches
# ----------------------------------------------------------------------------------------------------------------------
def rank_batch(cfg, rank, L=None, T=None, seed=0):
    """The synthetic batch of this rank: the LENGTHS / DURATIONS are those of seed `seed` on every rank (same valid-frame total
    per rank, so the weak-scaling value is not skewed by which rank drew the longest utterances); the CONTENT (symbols, energies,
    pitch, mel, speakers) is re-drawn per rank."""
    from daft_exprt_b200 import synthetic
    B, L, T = cfg['B'], L or cfg['L'], T or cfg['T']
    arrs = list(synthetic.make_batch(B, L, T, N_SPK_IDS, seed=seed, as_torch=False))
    if rank:
        rng = np.random.RandomState(7000 + 13 * rank + seed)
        symbols, dur_f, dur_i, s_en, s_pi, lens, f_en, f_pi, mel, out_lens, spk = arrs
        vl = np.arange(symbols.shape[1])[None, :] < lens[:, None]
        vt = np.arange(mel.shape[2])[None, :] < out_lens[:, None]
        live = vl & (dur_i > 0)
        arrs[0] = np.where(vl, rng.randint(1, 76, size=symbols.shape), 0).astype(np.int64)
        arrs[3] = np.where(live, rng.randn(*s_en.shape), 0).astype(np.float32)
        arrs[4] = np.where(live, rng.randn(*s_pi.shape), 0).astype(np.float32)
        arrs[6] = np.where(vt, rng.rand(*f_en.shape), 0).astype(np.float32)
        arrs[7] = np.where(vt & (rng.rand(*f_pi.shape) >= 0.3), 5.0 * rng.rand(*f_pi.shape), 0).astype(np.float32)
        arrs[8] = np.where(vt[:, None, :], np.clip(-5.0 + 2.0 * rng.randn(*mel.shape), -11.5, 2.0), 0).astype(np.float32)
        arrs[10] = rng.randint(0, N_SPK_IDS, size=spk.shape).astype(np.int64)
    return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in arrs)


def with_ids(tensors):
    B = tensors[0].shape[0]
    return tuple(tensors) + (['synthetic'] * B, [f'utt{i}' for i in range(B)])


def e2e_host_batches(cfg, rank):
    """Pinned FlatBatches of distinct (L_max, T_max), padded to bucket shapes (what a DataLoader with BucketedCollate yields)."""
    from daft_exprt_b200.data import BucketedCollate
    col = BucketedCollate(None, l_step=64, t_step=128, l_max=cfg['L'], t_max=cfg['T'])
    out = []
    # the stress config keeps ~26 GB of captured activations per bucket shape: two shapes there, four in the headline config
    shapes = E2E_SHAPES if cfg['B'] * cfg['T'] <= 64000 else (E2E_SHAPES[0], E2E_SHAPES[4])
    for k, (fl, ft) in enumerate(shapes):
        L, T = max(8, int(round(cfg['L'] * fl))), max(16, int(round(cfg['T'] * ft)))
        fb = col(with_ids(rank_batch(cfg, rank, L, T, seed=0)))   # same length profile as the fixed-shape batch, scaled to (L, T)
        out.append((fb.pin_memory() if torch.cuda.is_available() else fb, (L, T)))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def reference_objects(cfg, device='cpu'):
    import daft_exprt_oracle as oracle
    from daft_exprt_b200 import synthetic
    from daft_exprt_b200.model import reference_state_shapes
    sd = {k: v.to(device).requires_grad_(True) for k, v in synthetic.synthetic_state_dict(reference_state_shapes(N_SPK_IDS + 1), 1234).items()}
    return oracle, sd, oracle.OracleHParams(n_speakers=N_SPK_IDS + 1)


def reference_step(oracle, sd, ohp, inputs, train_step=True):
    """The reference algorithm (oracle port, fp32): forward + loss + backward."""
    targets = (inputs[1], inputs[3], inputs[4], inputs[8], inputs[10])
    if train_step:
        for v in sd.values():
            v.grad = None
        total, _ = oracle.loss(ohp, oracle.forward(sd, ohp, inputs), targets, 1000)
        total.backward()
    else:
        with torch.no_grad():
            total, _ = oracle.loss(ohp, oracle.forward(sd, ohp, inputs), targets, 1000)
    return total


def cpu_sample(cfg, sample_b=4):
    """Bounded sample of the workload: `sample_b` utterances SPREAD over the length-sorted rank-0 batch (first, last and two in
    between — not the longest ones), at their full lengths."""
    full = rank_batch(cfg, 0)
    B = cfg['B']
    idx = sorted({int(round(i * (B - 1) / (sample_b - 1))) for i in range(sample_b)})
    sub = tuple(t[idx].clone() for t in full)
    return sub, idx, int(sub[9].sum())


def cpu_reference_run(cfg, steps, warmup):
    threads = host_threads()
    torch.set_num_threads(threads)
    oracle, sd, ohp = reference_objects(cfg)
    sub, idx, frames = cpu_sample(cfg)
    for _ in range(warmup):
        reference_step(oracle, sd, ohp, sub)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step(oracle, sd, ohp, sub)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    sample = (f'utterances {idx} of the {cfg["B"]} of the length-sorted rank-0 batch (spread over the batch) at full L<={cfg["L"]}/'
              f'T<={cfg["T"]}, forward+loss+backward in fp32 (eval-mode math: the port has no dropout), {frames} valid frames per step, '
              f'{dt:.2f} s per step, torch.set_num_threads({threads})')
    return frames / dt, dt, {'value': frames / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample}


def cpu_baseline_subprocess(config_name):
    """The cpu_baseline leg of the `ours` arm: the reference arm itself in a FRESH process (clean OpenMP state: torchrun exports
    OMP_NUM_THREADS=1 and this process has long initialised its thread pools), all host threads."""
    env = {k: v for k, v in os.environ.items() if k not in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', config_name, '--steps', '2', '--warmup', '1'],
                         capture_output=True, text=True, env=env, timeout=900)
    try:
        return json.loads(out.stdout.strip().splitlines()[-1])['cpu_baseline']
    except Exception:
        return {'unavailable': (out.stderr or out.stdout)[-300:]}


def run_reference_arm(args, cfg):
    """--impl reference: the reference's algorithm on the host CPU (the Python reference cannot travel to the GPU box, so this
    is the pinned oracle port, all host threads), same metric/config, each step a bounded sample of the workload.  Under torchrun
    rank 0 alone runs; the other ranks exit 0."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    val, dt, base = cpu_reference_run(cfg, args.steps, args.warmup)
    print(json.dumps({
        'impl': 'reference', 'metric': cfg['metric'], 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'reference_arm': 'oracle port of the reference on host CPU (eval-mode forward + loss + backward, no optimiser step)'},
        'cpu_baseline': base,
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def gpu_eager_baseline(cfg, dev, steps=3, warmup=2):
    """Stock PyTorch eager on the same B200: the oracle port (plain torch ops: F.conv1d, explicit softmax attention, F.layer_norm,
    autograd) on cuda, fp32 with TF32 off, FULL rank-0 batch, forward + loss + backward + torch.optim.Adam step.  A baseline leg,
    like cpu_baseline: nothing on the product path touches it."""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    out = {}
    try:
        oracle, sd, ohp = reference_objects(cfg, dev)
        inputs = tuple(t.to(dev) for t in rank_batch(cfg, 0))
        frames = int(inputs[9].sum())
        opt = torch.optim.Adam(list(sd.values()), lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-6)
        for name, allow in (('fp32', False), ('tf32_allowed', True)):
            torch.backends.cuda.matmul.allow_tf32 = allow
            torch.backends.cudnn.allow_tf32 = allow

            def step():
                reference_step(oracle, sd, ohp, inputs)
                opt.step()
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'ms_per_step': ms, 'value': frames / (ms * 1e-3)}
        out.update(unit=UNIT, kind='oracle port in PyTorch eager on cuda:0 (F.conv1d / matmul / softmax / layer_norm + autograd + torch.optim.Adam), '
                              'eval-mode math, full batch', peak_memory_gb=torch.cuda.max_memory_allocated() / 1e9)
    except Exception as e:   # e.g. out of memory on the stress config: report, do not fail the bench
        out = {'unavailable': f'{type(e).__name__}: {e}'[:300]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out


# ----------------------------------------------------------------------------------------------------------------------
# kernel rooflines (rank 0): CUDA events around single launches on the launching stream, L2 flushed between launches
# ----------------------------------------------------------------------------------------------------------------------
def _time_launch(fn, flush, iters=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    k = max(1, len(ms) // 2)
    return sum(ms[:k]) / k * 1e-3


def time_dominant_kernel(cfg, dev, flush):
    """The FFT-block conv1 GEMM (B x T rows, 128 -> 1024 channels, k=3) through the C-ABI exactly as the training step issues it
    (operand planes in, bias + ReLU, hidden activation handed over as bf16 hi|lo planes)."""
    from daft_exprt_b200 import ops
    B, S, Cin, Cout, KW = cfg['B'], cfg['T'], 128, 1024, 3
    x = torch.randn(B, S, Cin, device=dev)
    w = torch.randn(Cout, Cin, KW, device=dev) * 0.05
    bias = torch.randn(Cout, device=dev)
    wp, _ = ops.packed(w)
    xP = ops.make_planes(x, B * S, Cin)
    t = _time_launch(lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP, emit_planes=True, want_y=False), flush)
    return 2.0 * B * S * Cout * Cin * KW, t


def time_attention_kernels(cfg, dev, flush, out_lens, p_drop):
    """dx_attention_fwd / dx_attention_bwd at the two head layouts of the model over the bench batch's own lengths.
    Algorithmic flops: forward 4 * len^2 * dh per (utterance, head) (QK^T + PV), backward 10 * len^2 * dh (five products)."""
    from daft_exprt_b200 import ops
    B, S = cfg['B'], cfg['T']
    lens = out_lens.to(dev)
    sq = float((out_lens.double() ** 2).sum())
    res = {}
    for name, H, dh, layers in (('prosody_encoder_8x16', 8, 16, 4), ('frame_decoder_2x64', 2, 64, 4)):
        D = H * dh
        qkv = torch.randn(B, S, 3 * D, device=dev)
        ctx = torch.empty(B, S, D, device=dev)
        lse = torch.empty(B, H, S, device=dev)
        planes = ops.attention_planes(B, S, H, dh, dev)
        ctxP = torch.empty(2, B * S, D, device=dev, dtype=torch.bfloat16)
        dctx = torch.randn(B, S, D, device=dev) * (torch.arange(S, device=dev)[None, :] < lens[:, None])[:, :, None]
        dqkv = torch.empty(B, S, 3 * D, device=dev)
        scratch = torch.empty(ops.lib().dx_attention_bwd_scratch_bytes(B, S, H, dh), device=dev, dtype=torch.uint8)
        fwd = lambda: ops._call('dx_attention_fwd', qkv.data_ptr(), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(), ops._p(planes),
                                ops._p(ctxP), B, S, H, dh, p_drop, 99, ops._st())
        bwd = lambda: ops._call('dx_attention_bwd', qkv.data_ptr(), ops._p(planes), lens.data_ptr(), ctx.data_ptr(), lse.data_ptr(),
                                dctx.data_ptr(), dqkv.data_ptr(), scratch.data_ptr(), B, S, H, dh, p_drop, 99, ops._st())
        tf, tb = _time_launch(fwd, flush), _time_launch(bwd, flush)
        res[name] = {'fwd_us': tf * 1e6, 'bwd_us': tb * 1e6, 'fwd_flops': 4.0 * sq * dh * H, 'bwd_flops': 10.0 * sq * dh * H, 'layers': layers}
    return res


# ----------------------------------------------------------------------------------------------------------------------
def run_inference_config(args, cfg, dev, rank, world, sampler):
    """configs[3]: DaftExprt.inference() at B=64 (per GPU), variable lengths; frames = generated mel frames."""
    from daft_exprt_b200 import cabi, synthetic
    from daft_exprt_b200.hparams import default_hparams
    from daft_exprt_b200.model import DaftExprt
    hp = default_hparams(n_speakers=N_SPK_IDS + 1)
    hp.stats = {f'spk {i}': {'pitch': {'mean': 5.0 + 0.05 * i, 'std': 0.25 + 0.01 * i}} for i in range(N_SPK_IDS)}
    torch.manual_seed(hp.seed)
    model = DaftExprt(hp).to(dev).eval()
    with torch.no_grad():   # durations ~ 58 ms per phoneme (5 frames at hop 256 / 22050 Hz): SURVEY.md section 8(d)
        lin = model.prosody_predictor.projection.linear_layer
        lin.weight.mul_(0.01)
        lin.bias.zero_()
        lin.bias[0] = 0.058
    host = tuple(t.pin_memory() for t in synthetic.make_inference_batch(cfg['B'], cfg['L'], cfg['T'], N_SPK_IDS, seed=3))
    resident = tuple(t.to(dev) for t in host)
    lib = cabi.load()

    pinned_out = {}

    def run(e2e):
        with torch.no_grad():
            inp = tuple(t.to(dev, non_blocking=True) for t in host) if e2e else resident
            enc, dec, _ = model.inference(inp, 'add', hp)
            if e2e:   # the generated mel-specs go back to pinned host memory (what generate.py hands to the vocoder / npz writer)
                key = tuple(dec[0].shape)
                if key not in pinned_out:
                    pinned_out[key] = torch.empty(key, dtype=torch.float32).pin_memory()
                pinned_out[key].copy_(dec[0], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return pinned_out[key], dec[1]
            return dec[0], dec[1]

    def timed(e2e, steps, warmup):
        for _ in range(warmup):
            run(e2e)
        torch.cuda.synchronize()
        l0 = lib.dx_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            mel, lens = run(e2e)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, int(lens.sum()), mel, (lib.dx_launch_count() - l0)
    ms, frames, mel, launches = timed(False, args.steps, max(args.warmup, 3))
    ms_e2e, _, mel_host, _ = timed(True, max(2, args.steps // 2), 1)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)
    h2d = sum(t.numel() * t.element_size() for t in host)
    line = {
        'metric': cfg['metric'], 'value': world * frames / (ms * 1e-3), 'unit': 'generated mel-frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x3 (bf16 hi/lo split operands on tcgen05, fp32 accumulate; fp32 state/LN/softmax)', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'global_batch': cfg['B'] * world, 'generated_frames_per_step': frames,
                   'audio_seconds_per_second': frames * 256 / 22050 / (ms * 1e-3), 'parallelism': f'replicas x{world}',
                   'launch': 'eager (one host read-back per batch for T_max, as the reference has)'},
        'e2e': {'value': world * frames / (ms_e2e * 1e-3), 'unit': 'generated mel-frames/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': mel_host.numel() * 4, 'ms_per_step': ms_e2e},
        'gpu_launches': int(launches), 'clocks': sampler.summary() if sampler else None,
    }
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='train', choices=sorted(CONFIGS))
    ap.add_argument('--backend', default='bf16x3', choices=['bf16x3', 'tf32', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true', help='skip the PyTorch-eager-on-GPU baseline leg')
    ap.add_argument('--eval', action='store_true', help='eval mode (dropout off)')
    ap.add_argument('--profile-step', action='store_true', help='bracket ONE extra step with cudaProfilerStart/Stop (for ncu --profile-from-start off)')
    ap.add_argument('--accumulation-steps', type=int, default=1, help='micro-batches per optimiser step (reference hparams.py:67 default: 3); a bench step = one micro-batch')
    ap.add_argument('--nccl-allreduce', action='store_true', help='N > 1: NCCL all-reduce + Adam instead of the fused reduce-scatter + Adam + all-gather kernel')
    ap.add_argument('--no-graph', action='store_true', help='issue the launches of a step eagerly instead of replaying the captured CUDA graph')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        return run_reference_arm(args, CONFIGS['train'] if cfg['kind'] != 'train' else cfg)

    import __graft_entry__ as entry
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if rank == 0:
        entry.build()
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local_rank))
        dist.barrier()
    if rank != 0:
        entry.build()
    from daft_exprt_b200 import cabi, ops
    from daft_exprt_b200.data import BatchPrefetcher
    from daft_exprt_b200.ddp import FlatAdam, FlatGradSync, FusedShardedAdam, broadcast_parameters
    from daft_exprt_b200.graph import GraphedTrainStep
    from daft_exprt_b200.hparams import default_hparams
    from daft_exprt_b200.loss import DaftExprtLoss, LossReadback
    from daft_exprt_b200.model import DaftExprt

    assert torch.cuda.is_available(), 'bench.py (impl=ours) needs a CUDA device: there is no CPU fallback'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    cabi.check(cabi.load().dx_device_check(), 'dx_device_check')
    ops.set_backend(args.backend)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if cfg['kind'] == 'infer':
        if sampler:
            sampler.start()
        run_inference_config(args, cfg, dev, rank, world, sampler)
        if world > 1:
            dist.destroy_process_group()
        return

    hp = default_hparams(n_speakers=N_SPK_IDS + 1)
    torch.manual_seed(hp.seed)
    model = DaftExprt(hp).to(dev)
    model.train(not args.eval)
    broadcast_parameters(model)
    crit = DaftExprtLoss(local_rank, hp)
    params = list(model.parameters())
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=hp.initial_learning_rate, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)

    # N > 1: the gradient exchange fused with the optimiser over NVLink peer memory (symmetric memory; NVSwitch multicast when mapped);
    # falls back to NCCL all-reduce + Adam when symmetric memory cannot be set up on this system (every rank takes the same branch)
    fused, exchange = None, 'none (one GPU)'
    if world > 1:
        exchange = 'NCCL all-reduce of the flat bucket + fused Adam'
        if not args.nccl_allreduce:
            ok = torch.ones(1, device=dev)
            try:
                fused = FusedShardedAdam(sync, opt)
            except Exception as e:   # noqa: BLE001
                ok.zero_()
                if rank == 0:
                    print(f'[bench] fused gradient exchange unavailable ({type(e).__name__}: {str(e)[:200]}); using NCCL all-reduce', file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                fused = None
            else:
                exchange = f'fused reduce-scatter + Adam + all-gather kernel over symmetric memory: {fused.mode}'

    host_batch = with_ids(tuple(t.pin_memory() for t in rank_batch(cfg, rank)))
    frames_rank = int(host_batch[9].sum())
    inputs, targets, _ = model.parse_batch(local_rank, host_batch)
    varied = e2e_host_batches(cfg, rank)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()

    graphed = None if args.no_graph else GraphedTrainStep(model, crit, sync, opt, fused_exchange=fused, accumulation_steps=args.accumulation_steps)

    def eager_step(inp, tgt, it, read_back=False):
        opt.zero_grad()
        out = crit.forward_device(model(inp), tgt, it)
        out[7].backward()
        if fused is not None:
            fused.step()
        else:
            sync.all_reduce_mean()
            opt.step()
        return out.tolist() if read_back else out

    def step_resident(it):
        if graphed is not None:   # same step, captured once and replayed as one CUDA graph (+ the NCCL all-reduce when N > 1)
            return graphed.step(inputs, targets, it)
        return eager_step(inputs, targets, it)

    prefetch = BatchPrefetcher(model, local_rank)   # the package's loader-side helper: parse_batch one step ahead on a side stream
    readback = LossReadback()                       # loss floats of step i collected while step i + 1 is in flight
    e2e_frames = [0]

    def make_e2e(batches):
        def step_e2e(it):
            # every step copies ITS OWN host batch (one pinned FlatBatch -> one H2D copy in parse_batch) and reads its loss back; the
            # copy of step i+1 is issued while step i computes, exactly as a training loop over a DataLoader would use the helper
            k = it % len(batches)
            if prefetch.pending is None:
                prefetch.submit(batches[k])
            inp, tgt, _ = prefetch.get()
            e2e_frames[0] += int(batches[k].tensors()[9].sum()) if hasattr(batches[k], 'tensors') else int(batches[k][9].sum())
            if graphed is not None:
                out = graphed.step(inp, tgt, it)                          # replay enqueued first ...
                prefetch.submit(batches[(it + 1) % len(batches)])         # ... then the next step's copy, while the GPU is busy
                return readback.submit(out)                               # ONE D2H copy of the 8 loss floats per step (values: one step late)
            prefetch.submit(batches[(it + 1) % len(batches)])
            return eager_step(inp, tgt, it, read_back=True)
        return step_e2e

    def timed(fn, steps, warmup, first_it=0):
        for i in range(warmup):
            fn(first_it + i)
        prefetch.drop()   # nothing copied before the timed region is consumed inside it
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e2e_frames[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = cabi.load().dx_launch_count() + (graphed.launches_replayed if graphed is not None else 0)
        e0.record()
        for i in range(steps):
            fn(first_it + warmup + i)
        readback.flush()   # the loss of the last e2e step is read inside the timed region too
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, cabi.load().dx_launch_count() + (graphed.launches_replayed if graphed is not None else 0) - l0

    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    # e2e, fixed shape (the same host batch every step)
    ms_e2e_fixed, _ = timed(make_e2e([host_batch]), max(2, args.steps // 2), 3)
    # e2e, varied shapes: warm-up covers every bucket shape once (graph captures), the timed region cycles all host batches
    nv = len(varied)
    h0, m0 = (graphed.hits, graphed.misses) if graphed is not None else (0, 0)
    steps_var = max(nv, args.steps // nv * nv)
    ms_e2e, _ = timed(make_e2e([fb for fb, _ in varied]), steps_var, nv)
    frames_var = torch.tensor([float(e2e_frames[0])], device=dev, dtype=torch.float64)
    hits, misses = (graphed.hits - h0, graphed.misses - m0) if graphed is not None else (0, 0)
    # the same step issued eagerly (no graph): the number a shape that was never captured pays
    ms_eager = None
    if graphed is not None and world == 1:
        g_keep, graphed = graphed, None
        ms_eager, _ = timed(step_resident, 3, 2)
        graphed = g_keep
    if args.profile_step:   # not timed: one step of the same workload for the ncu launch list
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step_resident(10 ** 6)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)
    peak_mem = torch.cuda.max_memory_allocated() / 1e9

    frames = torch.tensor([frames_rank], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(frames)
        dist.all_reduce(frames_var)
    total_frames = frames.item()
    h2d_var = sum(fb.nbytes for fb, _ in varied) / nv
    workload = cfg['workload'] if not args.eval else cfg['workload'].replace('train mode, dropout 0.1', 'eval mode')

    line = {
        'metric': cfg['metric'], 'value': total_frames / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'bf16x3': 'bf16x3 (bf16 hi/lo split operands on tcgen05, 3 passes, fp32 accumulate; weight-gradient sums over >= 4096 rows: 1 pass; '
                            'fp32 state/LN/softmax)',
                  'tf32': 'tf32', 'fp32': 'f32'}[args.backend],
        'data': 'synthetic',
        'config': {'workload': workload, 'global_batch': cfg['B'] * world, 'valid_frames_per_step': total_frames,
                   'padded_frames_per_step': cfg['B'] * world * cfg['T'], 'parallelism': f'dp{world}', 'grad_exchange': exchange,
                   'per_rank_batches': 'same lengths on every rank (equal valid-frame totals), per-rank content',
                   'accumulation_steps': args.accumulation_steps,
                   'launch': 'eager' if graphed is None else 'cuda-graph replay (graph.py); gpu_launches = kernels of libdaftexprt_b200.so executed by the replays',
                   'l2': 'per-step working set (several GB of activations) >> 126 MB L2; no explicit flush needed',
                   'peak_memory_gb': peak_mem},
        'e2e': {'value': frames_var.item() / steps_var / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d_var), 'd2h_bytes_per_step': 32,
                'ms_per_step': ms_e2e, 'valid_frames_per_step': frames_var.item() / steps_var,
                'shapes': [list(s) for _, s in varied], 'bucket_shapes': sorted({(fb.key()[0][1], fb.key()[8][2]) for fb, _ in varied}),
                'graph_hits': hits, 'graph_misses': misses, 'graph_hit_rate': hits / max(1, hits + misses),
                'note': 'every step takes a different pinned host batch (6 batches, bucket-padded FlatBatch, one H2D each) + one D2H copy of its 8 loss '
                        'floats (loss.LossReadback: values collected one step late, the last one before the clock stops); graph_hit_rate covers '
                        'warm-up (captures) + timed steps',
                'fixed_shape': {'value': total_frames / (ms_e2e_fixed * 1e-3), 'ms_per_step': ms_e2e_fixed,
                                'h2d_bytes_per_step': sum(t.numel() * t.element_size() for t in host_batch[:11])},
                'eager_no_graph_ms_per_step': ms_eager},
        'gpu_launches': int(launches),
    }
    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
        flops, t = time_dominant_kernel(cfg, dev, flush)
        passes = 3 if args.backend == 'bf16x3' else 1
        traffic = None
        for prof in ('r2_dominant_kernel.json', 'r1_dominant_kernel.json'):
            prof = os.path.join(ROOT, 'profiles', prof)
            if args.backend == 'bf16x3' and args.config == 'train' and os.path.exists(prof):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
                traffic = json.load(open(prof)).get('traffic_bytes_per_launch')
                break
        line['roofline'] = {'bound': 'tensor', 'achieved': flops / t / 1e12, 'peak': tf_burst, 'unit': 'TFLOP/s',
                            'frac': flops / t / 1e12 / tf_burst, 'traffic': traffic,
                            'kernel': f'gemm_tc_kernel (conv-GEMM, FFT-block conv1: {cfg["B"]}x{cfg["T"]} rows, 128->1024, k=3)',
                            'us_per_launch': t * 1e6,
                            'note': f'algorithmic flops 2*rows*Cout*Cin*KW per launch; peak = {src} bf16 dense burst; this backend issues '
                                    f'{passes} tensor-core pass(es) per algorithmic flop'}
        if args.backend != 'fp32':
            att = time_attention_kernels(cfg, dev, flush, host_batch[9], 0.0 if args.eval else 0.1)
            tot_f = sum((v['fwd_flops'] + v['bwd_flops']) * v['layers'] for v in att.values())
            tot_t = sum((v['fwd_us'] + v['bwd_us']) * v['layers'] for v in att.values()) * 1e-6
            # the phoneme encoder's 4 layers (L <= 200) are < 1 % of the attention flops and are left out
            line['roofline_attention'] = {
                'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': tf_burst, 'achieved': tot_f / tot_t / 1e12, 'frac': tot_f / tot_t / 1e12 / tf_burst,
                'attention_ms_per_step': tot_t * 1e3, 'attention_share_of_step': tot_t * 1e3 / ms_step,
                'kernels': {k: {'fwd_us': v['fwd_us'], 'bwd_us': v['bwd_us'], 'fwd_tflops': v['fwd_flops'] / v['fwd_us'] / 1e6,
                                'bwd_tflops': v['bwd_flops'] / v['bwd_us'] / 1e6, 'fwd_frac': v['fwd_flops'] / v['fwd_us'] / 1e6 / tf_burst,
                                'bwd_frac': v['bwd_flops'] / v['bwd_us'] / 1e6 / tf_burst} for k, v in att.items()},
                'note': 'algorithmic attention-GEMM flops (4 len^2 dh forward, 10 len^2 dh backward per head, valid lengths of the bench batch) / '
                        'time of dx_attention_fwd/bwd (incl. their operand-plane passes) over the 8 frame-side layers; 3 issued passes per flop'}
        if cfg['kind'] == 'train':
            # FLOPs of one step by the survey's formula (26.93 GFLOP forward per (L=200, T=1000) utterance, x3 for fwd+bwd)
            line['step_tflops_padded_formula'] = 3 * 26.93e9 * (cfg['T'] / 1000.0) * cfg['B'] * world / (ms_step * 1e-3) / 1e12
        del flush
        line['clocks'] = sampler.summary() if sampler else None
        if not args.no_eager_baseline and cfg['kind'] == 'train':
            if graphed is not None:
                graphed.cache.clear()          # release the captured graphs' private memory pools before the baseline leg
            torch.cuda.empty_cache()
            if cfg['B'] * cfg['T'] <= 64000:
                line['gpu_eager_baseline'] = gpu_eager_baseline(cfg, dev)
            else:   # explicit S x S attention probabilities of the eager port: 9.2 GB per tensor and layer at B=128, T=1500 (> 150 GB with autograd)
                line['gpu_eager_baseline'] = {'unavailable': 'skipped for this config: the eager port materialises B*H*T*T attention tensors (out of memory)'}
        if not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_subprocess('train' if args.config != 'stress' else 'stress')
        if args.config == 'stress':
            prof = os.path.join(ROOT, 'profiles', 'r2_stress_dram.json')
            if os.path.exists(prof):   # sum of dram__bytes_read + dram__bytes_write over the ncu launch list of ONE step of this config
                d = json.load(open(prof))
                gbs = d['dram_bytes_per_step'] / (ms_step * 1e-3) / 1e9
                line['hbm_roofline'] = {'bound': 'hbm', 'unit': 'GB/s', 'achieved': gbs, 'peak': hbm, 'frac': gbs / hbm,
                                        'dram_bytes_per_step': d['dram_bytes_per_step'], 'source': d.get('source')}
        print(json.dumps(line))
    if world > 1:
        # the other ranks wait for rank 0's CPU / eager legs on the rendezvous store (a blocking socket wait: an NCCL barrier would
        # spin one host core per rank and steal them from the CPU baseline)
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set('dx_bench_done', '1')
        else:
            store.wait(['dx_bench_done'])
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
