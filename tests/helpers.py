"""Shared helpers for the parity tests."""
import os
import zlib

import numpy as np
import torch

from daft_exprt_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GOLDEN_CASES = ('c1_1spk_b4', 'c2s_11spk_b6', 'ragged_11spk_b5')


def load_golden(case):
    return np.load(os.path.join(GOLDEN_DIR, case + '.npz'), allow_pickle=False)


def fake_stats(n_ids):
    return {f'spk {i}': {'pitch': {'mean': 5.0 + 0.05 * i, 'std': 0.25 + 0.01 * i}} for i in range(n_ids)}


def case_inputs(fx):
    n_ids, B, L, T, seed = (int(fx[k]) for k in ('meta_n_speaker_ids', 'meta_B', 'meta_L', 'meta_T', 'meta_batch_seed'))
    return synthetic.make_batch(B, L, T, n_ids, seed=seed), n_ids


def case_inference_inputs(fx, transform):
    n_ids, B, L, T, seed = (int(fx[k]) for k in ('meta_n_speaker_ids', 'meta_B', 'meta_L', 'meta_T', 'meta_batch_seed'))
    inp = list(synthetic.make_inference_batch(B, L, min(T, 300), n_ids, seed=seed))
    inp[1] = torch.from_numpy(fx['inf_dur_factors'])
    inp[2] = torch.from_numpy(fx['inf_energy_factors'])
    inp[3] = torch.from_numpy(fx[f'inf_{transform}_pitch_factors'])
    return tuple(inp)


def targets_of(inputs):
    # model.py:750 — targets = (durations_float, symbols_energy, symbols_pitch, mel_specs, speaker_ids)
    return (inputs[1], inputs[3], inputs[4], inputs[8], inputs[10])


def grad_projection(name, g):
    rng = np.random.RandomState(zlib.crc32(('proj:' + name).encode()) & 0x7fffffff)
    v = rng.randn(g.numel()).astype(np.float64)
    return float(np.dot(g.detach().double().cpu().numpy().ravel(), v))


def scale_rel_err(a, b):
    """max|a-b| / max|b|  — error relative to the tensor's own scale (the parity metric for dense fp tensors)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / max(denom, 1e-30)


def l2_rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return (a - b).norm().item() / max(b.norm().item(), 1e-30)
