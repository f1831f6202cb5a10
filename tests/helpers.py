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


# ----------------------------------------------------------------------------------------------------------------------
# the CUDA path's counter-based dropout masks, restated in numpy (csrc/common.cuh: hash_u32 / hash_uniform / dropout_scale /
# drop_keep / drop_threshold) so that the oracle can replay a train-mode step with EXACTLY the masks the kernels used
# ----------------------------------------------------------------------------------------------------------------------
def hash_u32(seed, idx):
    """common.cuh hash_u32(seed, idx): idx = uint64 numpy array, seed = Python int (64 bit) -> uint32 array."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    idx = np.asarray(idx, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    x = ((idx & m32) * np.uint64(0x9E3779B1) + np.uint64(seed & 0xFFFFFFFF)) & m32
    x ^= ((idx >> np.uint64(32)) * np.uint64(0x85EBCA77) + np.uint64(seed >> 32)) & m32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & m32
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & m32
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def rows_dropout_scale(seed, n, p):
    """dropout_scale over element indices 0..n-1 (norm.cu: idx = row * D + c): float32 array of {0, 1/(1-p)}."""
    u = (hash_u32(seed, np.arange(n, dtype=np.uint64)) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    inv_keep = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(u >= np.float32(p), inv_keep, np.float32(0.0)).astype(np.float32)


def attn_dropout_scale(seed, B, H, S, p):
    """attention_tc.cu / attention_bwd_tc.cu: row key = hash_u32(seed, (b*H + h)*S + q); keep <=> mix(rk ^ k*0x9E3779B1) >= p*2^32."""
    rk = hash_u32(seed, np.arange(B * H * S, dtype=np.uint64)).reshape(B, H, S, 1)
    col = (np.arange(S, dtype=np.uint64) * np.uint64(0x9E3779B1) & np.uint64(0xFFFFFFFF)).astype(np.uint32).reshape(1, 1, 1, S)
    x = rk ^ col
    x = x ^ (x >> np.uint32(16))
    x = (x.astype(np.uint64) * np.uint64(0x85EBCA6B) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    thresh = np.uint32(int(np.float32(min(p, 0.99999994)) * np.float32(4294967296.0)))
    inv_keep = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(x >= thresh, inv_keep, np.float32(0.0)).astype(np.float32)


class DropoutReplay:
    """Dropout hook for `oracle.forward(..., dropout=...)`: consumes the seeds the CUDA path drew (`ops.next_seed`, recorded by
    `record_seeds`) in the same call order and applies the identical masks."""

    def __init__(self, seeds):
        self.seeds, self.k = list(seeds), 0

    def _next(self):
        s = self.seeds[self.k]
        self.k += 1
        return s

    def rows(self, x, p):
        scale = rows_dropout_scale(self._next(), x.numel(), p).reshape(tuple(x.shape))
        return x * torch.from_numpy(scale).to(x.dtype)

    def attn(self, probs, p):
        B, H, S, _ = probs.shape
        return probs * torch.from_numpy(attn_dropout_scale(self._next(), B, H, S, p)).to(probs.dtype)


class record_seeds:
    """Context manager: records every dropout seed `ops.next_seed()` hands out (in call order)."""

    def __enter__(self):
        from daft_exprt_b200 import ops
        self.ops, self.orig, self.seeds = ops, ops.next_seed, []

        def wrapped():
            s = self.orig()
            self.seeds.append(s)
            return s
        ops.next_seed = wrapped
        return self.seeds

    def __exit__(self, *exc):
        self.ops.next_seed = self.orig
