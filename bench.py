#!/usr/bin/env python
"""Headline benchmark: valid mel-frames/s of the Daft-Exprt training step (forward + loss + backward + gradient all-reduce +
fused Adam) on synthetic batches of BASELINE.json configs[1]/[2] (11-speaker hparams, B=32 per GPU, L<=200, T<=1000, 80 mels).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's sm_100a path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps 1 --warmup 0          # the reference algorithm on the host CPU (oracle port)

Prints ONE JSON line (rank 0).  `value` = whole-job valid frames/s with inputs resident in HBM; `e2e` = the same metric through
the public module API (`parse_batch` from pinned host memory every step + loss read-back); `roofline` = the dominant kernel
(tcgen05 conv-GEMM) timed live with CUDA events; `cpu_baseline` = the oracle port on this box's host cores (bounded sample).
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

import torch  # noqa: E402

METRIC = 'mel_frames_per_sec_train_step'
UNIT = 'valid mel-frames/s'
B_PER_GPU, L_MAX, T_MAX, N_SPK_IDS = 32, 200, 1000, 11
WORKLOAD = ('configs[1]/[2]: 11-speaker LJ+ESD hparams, B=32 per GPU, L<=200 phonemes, T<=1000 frames, 80 mels; step = forward + '
            'DaftExprtLoss + backward + flat-bucket grad all-reduce (N>1) + fused Adam; train mode, dropout 0.1')


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


def make_host_batch(rank, seed=0):
    from daft_exprt_b200 import synthetic
    inputs = synthetic.make_batch(B_PER_GPU, L_MAX, T_MAX, N_SPK_IDS, seed=1000 * seed + rank)
    dirs, files = ['synthetic'] * B_PER_GPU, [f'utt{i}' for i in range(B_PER_GPU)]
    return tuple(t.pin_memory() if torch.cuda.is_available() else t for t in inputs) + (dirs, files)


def cpu_reference_step(inputs, sd, ohp, train_step=True):
    """The reference algorithm (oracle port, fp32, all host threads): forward + loss (+ backward)."""
    import daft_exprt_oracle as oracle
    targets = (inputs[1], inputs[3], inputs[4], inputs[8], inputs[10])
    t0 = time.perf_counter()
    if train_step:
        for v in sd.values():
            v.grad = None
        total, _ = oracle.loss(ohp, oracle.forward(sd, ohp, inputs), targets, 1000)
        total.backward()
    else:
        with torch.no_grad():
            oracle.loss(ohp, oracle.forward(sd, ohp, inputs), targets, 1000)
    return time.perf_counter() - t0


def cpu_baseline(sample_b=4, steps=1):
    """Bounded CPU sample of the same workload: the first `sample_b` utterances of the rank-0 batch (full L/T), fwd+loss+bwd."""
    import daft_exprt_oracle as oracle
    from daft_exprt_b200 import synthetic
    from daft_exprt_b200.model import reference_state_shapes
    inputs = synthetic.make_batch(B_PER_GPU, L_MAX, T_MAX, N_SPK_IDS, seed=0)
    sub = tuple(t[:sample_b].clone() for t in inputs)
    sd = {k: v.requires_grad_(True) for k, v in synthetic.synthetic_state_dict(reference_state_shapes(N_SPK_IDS + 1), 1234).items()}
    ohp = oracle.OracleHParams(n_speakers=N_SPK_IDS + 1)
    frames = int(sub[9].sum())
    times = [cpu_reference_step(sub, sd, ohp) for _ in range(steps)]
    t = min(times)
    return {'value': frames / t, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'first {sample_b} of the {B_PER_GPU} utterances of the rank-0 batch at full L<=200/T<=1000 (eval-mode math: the '
                      f'oracle has no dropout), forward+loss+backward in fp32, {frames} valid frames in {t:.2f} s'}


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on the host CPU (the Python reference cannot travel to the GPU box, so this
    is the pinned oracle port, all host threads), same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import daft_exprt_oracle as oracle
    from daft_exprt_b200 import synthetic
    from daft_exprt_b200.model import reference_state_shapes
    sample_b = 4
    inputs = synthetic.make_batch(B_PER_GPU, L_MAX, T_MAX, N_SPK_IDS, seed=0)
    sub = tuple(t[:sample_b].clone() for t in inputs)
    sd = {k: v.requires_grad_(True) for k, v in synthetic.synthetic_state_dict(reference_state_shapes(N_SPK_IDS + 1), 1234).items()}
    ohp = oracle.OracleHParams(n_speakers=N_SPK_IDS + 1)
    frames = int(sub[9].sum())
    for _ in range(args.warmup):
        cpu_reference_step(sub, sd, ohp)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sub, sd, ohp)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = frames / dt
    sample = f'{sample_b} of {B_PER_GPU} utterances per step at full L<=200/T<=1000, forward+loss+backward, fp32, {frames} valid frames/step'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'reference_arm': 'oracle port of the reference on host CPU'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def time_dominant_kernel(dev, iters=20):
    """Roofline of the dominant kernel: the FFT-block conv1 GEMM (32x1000 rows, 128 -> 1024 channels, k=3) through the
    C-ABI exactly as the training step issues it (operand planes in, bias + ReLU, hidden activation handed over as bf16
    hi|lo planes), CUDA events around each launch on the launching stream, L2 flushed between launches."""
    from daft_exprt_b200 import ops
    B, S, Cin, Cout, KW = B_PER_GPU, T_MAX, 128, 1024, 3
    x = torch.randn(B, S, Cin, device=dev)
    w = torch.randn(Cout, Cin, KW, device=dev) * 0.05
    bias = torch.randn(Cout, device=dev)
    wp, _ = ops.packed(w)
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    xP = ops.make_planes(x, B * S, Cin)
    run = lambda: ops.conv_gemm(None, wp, bias, B, S, relu=True, x_planes=xP, emit_planes=True, want_y=False)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    # enqueue everything first (the 512 MB flush kernels give the host time to run ahead), then read the events:
    # e0 -> e1 brackets exactly one launch of gemm_tc_kernel
    evs = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    t = sum(ms[: max(1, len(ms) // 2)]) / max(1, len(ms) // 2) * 1e-3
    flops = 2.0 * B * S * Cout * Cin * KW
    return flops, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--backend', default='bf16x3', choices=['bf16x3', 'tf32', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eval', action='store_true', help='eval mode (dropout off)')
    ap.add_argument('--profile-step', action='store_true', help='bracket ONE extra step with cudaProfilerStart/Stop (for ncu --profile-from-start off)')
    ap.add_argument('--no-graph', action='store_true', help='issue the ~640 launches of a step eagerly instead of replaying the captured CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

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
    from daft_exprt_b200 import cabi, ops, synthetic
    from daft_exprt_b200.data import BatchPrefetcher
    from daft_exprt_b200.ddp import FlatAdam, FlatGradSync, broadcast_parameters
    from daft_exprt_b200.graph import GraphedTrainStep
    from daft_exprt_b200.hparams import default_hparams
    from daft_exprt_b200.loss import DaftExprtLoss
    from daft_exprt_b200.model import DaftExprt

    assert torch.cuda.is_available(), 'bench.py (impl=ours) needs a CUDA device: there is no CPU fallback'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    cabi.check(cabi.load().dx_device_check(), 'dx_device_check')
    ops.set_backend(args.backend)

    hp = default_hparams(n_speakers=N_SPK_IDS + 1)
    torch.manual_seed(hp.seed)
    model = DaftExprt(hp).to(dev)
    model.train(not args.eval)
    broadcast_parameters(model)
    crit = DaftExprtLoss(local_rank, hp)
    params = list(model.parameters())
    sync = FlatGradSync(params, mode='gather')
    opt = FlatAdam(params, sync, lr=hp.initial_learning_rate, betas=hp.betas, eps=hp.epsilon, weight_decay=hp.weight_decay)

    host_batch = make_host_batch(rank)
    frames_rank = int(host_batch[9].sum())
    inputs, targets, _ = model.parse_batch(local_rank, host_batch)
    torch.cuda.synchronize()

    graphed = None if args.no_graph else GraphedTrainStep(model, crit, sync, opt)

    def step_resident(it):
        if graphed is not None:   # same step, captured once and replayed as one CUDA graph (+ the NCCL all-reduce when N > 1)
            return graphed.step(inputs, targets, it)
        opt.zero_grad()
        out = crit.forward_device(model(inputs), targets, it)
        out[7].backward()
        sync.all_reduce_mean()
        opt.step()
        return out

    prefetch = BatchPrefetcher(model, local_rank)   # the package's loader-side helper: parse_batch one step ahead on a side stream

    def step_e2e(it):
        # every step copies ITS inputs from pinned host memory (11 H2D copies = model.parse_batch) and reads its loss back; the
        # copies of step i+1 are issued while step i computes, exactly as a training loop over a DataLoader would use the helper
        if prefetch.pending is None:
            prefetch.submit(host_batch)
        inp, tgt, _ = prefetch.get()
        if graphed is not None:
            out = graphed.step(inp, tgt, it)                          # replay enqueued first ...
            prefetch.submit(host_batch)                               # ... then the next step's copies, while the GPU is busy
            return out.tolist()                                       # ONE D2H read of the 8 loss floats
        prefetch.submit(host_batch)
        opt.zero_grad()
        loss, terms = crit(model(inp), tgt, it)                        # ONE D2H read of the 8 loss floats
        loss.backward()
        sync.all_reduce_mean()
        opt.step()
        return terms

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        prefetch.drop()   # nothing copied before the timed region is consumed inside it
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = cabi.load().dx_launch_count() + (graphed.launches_replayed if graphed is not None else 0)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, cabi.load().dx_launch_count() + (graphed.launches_replayed if graphed is not None else 0) - l0

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, max(args.warmup, 3))
    ms_e2e, _ = timed(step_e2e, max(2, args.steps // 2), 1)
    if args.profile_step:   # not timed: one step of the same workload for the ncu launch list
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step_resident(10 ** 6)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)

    frames = torch.tensor([frames_rank], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(frames)
    total_frames = frames.item()
    h2d = sum(t.numel() * t.element_size() for t in host_batch[:11])

    line = {
        'metric': METRIC, 'value': total_frames / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'bf16x3': 'bf16x3 (bf16 hi/lo split operands on tcgen05, fp32 accumulate; fp32 state/LN/softmax)',
                  'tf32': 'tf32', 'fp32': 'f32'}[args.backend],
        'data': 'synthetic',
        'config': {'workload': WORKLOAD if not args.eval else WORKLOAD.replace('train mode, dropout 0.1', 'eval mode'),
                   'global_batch': B_PER_GPU * world, 'valid_frames_per_step': total_frames,
                   'padded_frames_per_step': B_PER_GPU * world * T_MAX, 'parallelism': f'dp{world}',
                   'launch': 'eager' if graphed is None else 'cuda-graph replay (graph.py); gpu_launches = kernels of libdaftexprt_b200.so executed by the replays',
                   'l2': 'per-step working set (several GB of activations) >> 126 MB L2; no explicit flush needed'},
        'e2e': {'value': total_frames / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 32,
                'ms_per_step': ms_e2e},
        'gpu_launches': int(launches),
    }
    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        flops, t = time_dominant_kernel(dev)
        passes = 3 if args.backend == 'bf16x3' else 1
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'r1_dominant_kernel.json')
        if args.backend == 'bf16x3' and os.path.exists(prof):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
            traffic = json.load(open(prof)).get('traffic_bytes_per_launch')
        line['roofline'] = {'bound': 'tensor', 'achieved': flops / t / 1e12, 'peak': tf_burst, 'unit': 'TFLOP/s',
                            'frac': flops / t / 1e12 / tf_burst, 'traffic': traffic,
                            'kernel': 'gemm_tc_kernel (conv-GEMM, FFT-block conv1: 32x1000 rows, 128->1024, k=3)',
                            'note': f'algorithmic flops 2*rows*Cout*Cin*KW per launch; peak = {src} bf16 dense burst; this backend issues '
                                    f'{passes} tensor-core pass(es) per algorithmic flop'}
        line['clocks'] = sampler.summary() if sampler else None
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
