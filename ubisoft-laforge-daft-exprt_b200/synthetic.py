"""Seeded synthetic batches and weights for the Daft-Exprt mel-prediction path.

Everything here is numpy-`RandomState` based so that the same seed gives bit-identical inputs and weights in
the authoring container (where the golden vectors are generated from the real reference) and on the GPU box
(where `/root/reference` does not exist), independently of torch's RNG implementation.

Batch tuple layout = what the reference's collate produces and `DaftExprt.parse_batch` consumes
(reference `data_loader.py:140-211`, `model.py:727-753`): right-zero-padded, sorted by decreasing input length,
`sum(durations_int[b]) == output_lengths[b]`, zero-duration symbols are common (word boundaries / EOS).
"""
import zlib

import numpy as np
import torch

HOP_LENGTH = 256
SAMPLING_RATE = 22050


def make_batch(batch_size, max_symbols, max_frames, n_speaker_ids, seed=0, n_mels=80, min_len_frac=0.5,
               zero_dur_prob=0.15, as_torch=True):
    """Return the 11-tuple `inputs` of `DaftExprt.forward` (model.py:759-760) as CPU tensors.

    lengths ~ U{L/2..L} sorted desc with the first == L; durations_int ~ U{1..d_max} with >= 15 % zeros and the
    row sum clipped to <= max_frames; the longest row is forced to sum to exactly max_frames.
    """
    rng = np.random.RandomState(seed)
    B, L, T = batch_size, max_symbols, max_frames
    lens = rng.randint(max(1, int(L * min_len_frac)), L + 1, size=B)
    lens[0] = L
    lens = np.sort(lens)[::-1].copy()
    d_max = max(2, int(round(2.0 * T / (L * (1.0 - zero_dur_prob)))) - 1)

    symbols = np.zeros((B, L), np.int64)
    dur_int = np.zeros((B, L), np.int64)
    for b in range(B):
        n = int(lens[b])
        symbols[b, :n] = rng.randint(1, 76, size=n)
        d = rng.randint(1, d_max + 1, size=n)
        d[rng.rand(n) < zero_dur_prob] = 0
        if d.sum() == 0:
            d[0] = 1
        # clip the row total to <= T by trimming from the end
        over = int(d.sum()) - T
        i = n - 1
        while over > 0 and i >= 0:
            take = min(over, int(d[i]))
            d[i] -= take
            over -= take
            i -= 1
        dur_int[b, :n] = d
    # force the first (longest-input) row to define T_max == T exactly: spread the shortfall over its voiced symbols
    short = T - int(dur_int[0].sum())
    nz = np.nonzero(dur_int[0, :lens[0]])[0]
    if len(nz) == 0:
        nz = np.array([0])
    dur_int[0, nz] += short // len(nz)
    dur_int[0, nz[:short % len(nz)]] += 1
    out_lens = dur_int.sum(axis=1)
    # the reference requires T_max = max(output_lengths); keep the batch sorted by input length only
    dur_float = (dur_int.astype(np.float64) * HOP_LENGTH / SAMPLING_RATE).astype(np.float32)

    sym_energy = rng.randn(B, L).astype(np.float32)
    sym_pitch = rng.randn(B, L).astype(np.float32)
    dead = dur_int == 0
    sym_energy[dead] = 0.
    sym_pitch[dead] = 0.

    Tm = int(out_lens.max())
    frames_energy = rng.rand(B, Tm).astype(np.float32)
    frames_pitch = (5.0 * rng.rand(B, Tm)).astype(np.float32)
    frames_pitch[rng.rand(B, Tm) < 0.3] = 0.
    mel = np.clip(-5.0 + 2.0 * rng.randn(B, n_mels, Tm), -11.5, 2.0).astype(np.float32)
    for b in range(B):
        frames_energy[b, out_lens[b]:] = 0.
        frames_pitch[b, out_lens[b]:] = 0.
        mel[b, :, out_lens[b]:] = 0.
    speaker_ids = rng.randint(0, n_speaker_ids, size=B).astype(np.int64)

    arrs = (symbols, dur_float, dur_int, sym_energy, sym_pitch, lens.astype(np.int64),
            frames_energy, frames_pitch, mel, out_lens.astype(np.int64), speaker_ids)
    if not as_torch:
        return arrs
    return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in arrs)


def make_inference_batch(batch_size, max_symbols, max_ref_frames, n_speaker_ids, seed=0, n_mels=80):
    """Return the 10-tuple `inputs` of `DaftExprt.inference` (model.py:879-880) as CPU tensors."""
    rng = np.random.RandomState(seed + 7919)
    B, L, T = batch_size, max_symbols, max_ref_frames
    lens = rng.randint(max(1, L // 2), L + 1, size=B)
    lens[0] = L
    lens = np.sort(lens)[::-1].copy()
    symbols = np.zeros((B, L), np.int64)
    for b in range(B):
        symbols[b, :lens[b]] = rng.randint(1, 76, size=int(lens[b]))
    dur_factors = np.ones((B, L), np.float32)
    energy_factors = np.ones((B, L), np.float32)
    pitch_factors = np.zeros((B, L), np.float32)
    ref_lens = rng.randint(max(1, T // 2), T + 1, size=B).astype(np.int64)
    ref_lens[rng.randint(0, B)] = T
    energy_refs = rng.rand(B, T).astype(np.float32)
    pitch_refs = (5.0 * rng.rand(B, T)).astype(np.float32)
    pitch_refs[rng.rand(B, T) < 0.3] = 0.
    mel_refs = np.clip(-5.0 + 2.0 * rng.randn(B, n_mels, T), -11.5, 2.0).astype(np.float32)
    for b in range(B):
        energy_refs[b, ref_lens[b]:] = 0.
        pitch_refs[b, ref_lens[b]:] = 0.
        mel_refs[b, :, ref_lens[b]:] = 0.
    speaker_ids = rng.randint(0, n_speaker_ids, size=B).astype(np.int64)
    arrs = (symbols, dur_factors, energy_factors, pitch_factors, lens.astype(np.int64),
            energy_refs, pitch_refs, mel_refs, ref_lens, speaker_ids)
    return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in arrs)


def synthetic_state_dict(shapes, seed=1234):
    """Deterministic weights keyed by parameter NAME (not by construction order).

    `shapes` is `{name: shape}` (e.g. from `model.state_dict()`).  Matrices/conv kernels get a Xavier-like scale,
    LayerNorm weights are 1 + small noise, biases are small but non-zero (so that halo/padding effects driven by
    biases are exercised).
    """
    out = {}
    for name, shape in shapes.items():
        shape = tuple(int(s) for s in shape)
        rng = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7fffffff)
        if name.endswith('post_multipliers'):
            a = 0.5 + 0.5 * rng.rand(*shape)
        elif len(shape) >= 2:
            fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
            fan_in = int(np.prod(shape[1:]))
            a = rng.randn(*shape) * np.sqrt(2.0 / (fan_in + fan_out))
            if 'embedding.weight' in name:
                a = rng.randn(*shape) * 0.3
        elif 'layer_norm' in name or _is_seq_layernorm(name):
            a = (1.0 + 0.05 * rng.randn(*shape)) if name.endswith('weight') else 0.05 * rng.randn(*shape)
        else:
            a = 0.05 * rng.randn(*shape)
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def _is_seq_layernorm(name):
    # LayerNorms living inside nn.Sequential: prosody_encoder.convs.{2,6,10}, prosody_predictor.blocks.0.{2,6}
    parts = name.split('.')
    if parts[0] == 'prosody_encoder' and parts[1] == 'convs' and parts[2] in ('2', '6', '10'):
        return True
    if parts[0] == 'prosody_predictor' and parts[1] == 'blocks' and parts[3] in ('2', '6'):
        return True
    return False


def nudge_for_inference(state_dict, bias=0.07, scale=0.02):
    """Shrink the duration head so that predicted durations land in ~[0.03, 0.11] s;
    with raw synthetic weights the reference overflows its 5000-row positional table (model.py:123,147)."""
    sd = dict(state_dict)
    w = sd['prosody_predictor.projection.linear_layer.weight'].clone()
    b = sd['prosody_predictor.projection.linear_layer.bias'].clone()
    w[0] *= scale
    b[0] = bias
    sd['prosody_predictor.projection.linear_layer.weight'] = w
    sd['prosody_predictor.projection.linear_layer.bias'] = b
    return sd
