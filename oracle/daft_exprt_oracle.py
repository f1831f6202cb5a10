"""ORACLE — plain PyTorch fp32/fp64 CPU restatement of the Daft-Exprt mel-prediction path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product path
(`ubisoft-laforge-daft-exprt_b200/`) never does and fails loudly when its CUDA extension is missing.

Why a restatement exists at all: the reference is pure Python/PyTorch and its arithmetic lives in torch
(`nn.MultiheadAttention`, `nn.Conv1d`, ... pinned torch==1.9.0, `setup.py:25`), it needs three harness shims to
even import and `/root/reference` does not travel to the GPU box.  This restatement is functional
(a flat `state_dict` in, tensors out), uses no `nn.Module`, spells attention out explicitly, and is PINNED against
the real reference: `tests/golden/make_golden.py` runs the unmodified reference module in the authoring container
and commits its outputs under `tests/golden/`; `tests/test_oracle_vs_golden.py` checks this file against them
(and against the live reference whenever `/root/reference` is present).  Parity status: PINNED by those fixtures
(the reference itself ships no tests or golden vectors).

Every function cites the reference lines it restates (paths relative to `/root/reference/src/daft_exprt/`).
Eval-mode semantics by default (all dropouts are identity), which is what the golden fixtures pin.  Train-mode
parity: `forward(..., dropout=hook)` applies a caller-supplied mask at every `nn.Dropout` / attention-dropout site of the
reference, in call order (`hook.rows(x, p)` for element-wise sites, `hook.attn(probs, p)` for the attention weights); the tests
replay the CUDA path's own counter-based masks through it (tests/helpers.py: DropoutReplay), so one train-mode step can be
compared numerically instead of statistically.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------------
# hyper-parameters that the path reads (hparams.py:44,71,73,90-128,189,200-201; loss.py:9-16)
# ----------------------------------------------------------------------------------------------------------------------
class OracleHParams:
    def __init__(self, n_speakers=12, n_symbols=76, n_mel_channels=80, stats=None, **kw):
        self.n_speakers = n_speakers
        self.n_symbols = n_symbols
        self.n_mel_channels = n_mel_channels
        self.lambda_reversal = 1.
        self.post_mult_weight = 1e-3
        self.adv_max_weight = 1e-2
        self.warmup_steps = 10000
        self.dur_weight = self.energy_weight = self.pitch_weight = self.mel_spec_weight = 1.
        self.filter_length, self.hop_length, self.sampling_rate, self.centered = 1024, 256, 22050, True
        # dropout rates: hparams.py:90-128 (only used when a dropout hook is passed to forward)
        self.prosody_encoder = dict(nb_blocks=4, hidden_embed_dim=128, attn_nb_heads=8, conv_kernel=3, conv_channels=1024,
                                    attn_dropout=0.1, conv_dropout=0.1)
        self.phoneme_encoder = dict(nb_blocks=4, hidden_embed_dim=128, attn_nb_heads=2, conv_kernel=3, conv_channels=1024,
                                    attn_dropout=0.1, conv_dropout=0.1)
        self.local_prosody_predictor = dict(nb_blocks=1, conv_kernel=3, conv_channels=256, conv_dropout=0.1)
        self.gaussian_upsampling_module = dict(conv_kernel=3)
        self.frame_decoder = dict(nb_blocks=4, attn_nb_heads=2, conv_kernel=3, conv_channels=1024, attn_dropout=0.1, conv_dropout=0.1)
        self.stats = stats or {}
        for k, v in kw.items():
            setattr(self, k, v)


# ----------------------------------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------------------------------
def valid_mask(lengths, max_len=None):
    """model.py:14-24 — True where position < length.  (B,) int64 -> (B, max_len) bool."""
    max_len = int(lengths.max()) if max_len is None else max_len
    return torch.arange(max_len, device=lengths.device)[None, :] < lengths[:, None]


def positional_table(max_len, dim, dtype=torch.float32, timestep=10000.):
    """model.py:123-130 — pe[p, 2i] = sin(p * exp(-2i ln(1e4)/D)), pe[p, 2i+1] = cos(same)."""
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2).float() * (-np.log(timestep) / dim))
    pe = torch.zeros(max_len, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def positional_encoding(lengths, max_len, dim, dtype):
    """model.py:132-150 called with N=1 (model.py:400,499,696) => absolute positions 0..len-1, zeros beyond."""
    pe = positional_table(max_len, dim, dtype).to(lengths.device)
    return pe[None, :, :] * valid_mask(lengths, max_len)[:, :, None].to(dtype)


def conv1d_cl(x, w, b, relu=False):
    """ConvNorm1D, model.py:86-94 — channels-last Conv1d, stride 1, 'same' zero padding over the PADDED layout
    (no masking between convs: the halo leak of SURVEY.md §0.6 is part of the reference's result)."""
    k = w.shape[2]
    y = F.conv1d(x.transpose(1, 2), w, b, padding=(k - 1) // 2).transpose(1, 2)
    return F.relu(y) if relu else y


def self_attention(x, lengths, sd, prefix, nb_heads, dropout=None, p=0.):
    """MultiHeadAttention, model.py:171-193 around nn.MultiheadAttention (model.py:165, math path):
    q,k,v = split(x W_in^T + b_in); scores = (q/sqrt(dh)) k^T with key padding -> -inf; softmax; PV; out-proj;
    then LN(dropout(attn) + x) (model.py:189-191); nn.MultiheadAttention drops the softmax weights with the same p (model.py:165)."""
    B, S, D = x.shape
    dh = D // nb_heads
    w_in, b_in = sd[prefix + 'multi_head_attention.in_proj_weight'], sd[prefix + 'multi_head_attention.in_proj_bias']
    qkv = F.linear(x, w_in, b_in)
    q, k, v = qkv.split(D, dim=2)

    def heads(t):
        return t.reshape(B, S, nb_heads, dh).permute(0, 2, 1, 3)
    q, k, v = heads(q) * (1.0 / math.sqrt(dh)), heads(k), heads(v)
    scores = q @ k.transpose(-1, -2)  # (B, H, S, S)
    key_ok = valid_mask(lengths, S)[:, None, None, :]
    scores = scores.masked_fill(~key_ok, float('-inf'))
    probs = torch.softmax(scores, dim=-1)
    if dropout is not None:
        probs = dropout.attn(probs, p)
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, S, D)
    out = F.linear(ctx, sd[prefix + 'multi_head_attention.out_proj.weight'], sd[prefix + 'multi_head_attention.out_proj.bias'])
    if dropout is not None:
        out = dropout.rows(out, p)
    return F.layer_norm(out + x, (D,), sd[prefix + 'layer_norm.weight'], sd[prefix + 'layer_norm.bias'])


def conv_feed_forward(x, film, sd, prefix, dropout=None, p=0.):
    """PositionWiseConvFF, model.py:220-237 — gamma * LN(conv2(relu(conv1(x))) + x) + beta."""
    D = x.shape[2]
    h = conv1d_cl(x, sd[prefix + 'convs.0.conv.weight'], sd[prefix + 'convs.0.conv.bias'], relu=True)
    y = conv1d_cl(h, sd[prefix + 'convs.2.conv.weight'], sd[prefix + 'convs.2.conv.bias'])
    if dropout is not None:   # model.py:213-217: the Sequential ends with nn.Dropout
        y = dropout.rows(y, p)
    y = F.layer_norm(y + x, (D,), sd[prefix + 'layer_norm.weight'], sd[prefix + 'layer_norm.bias'])
    if film is not None:
        assert film.shape[1] == 2 * D
        y = film[:, None, :D] * y + film[:, None, D:]
    return y


def fft_block(x, film, lengths, sd, prefix, nb_heads, dropout=None, cfg=None):
    """FFTBlock, model.py:251-264 — attention, zero padded rows, conv-FF (+FiLM), zero padded rows."""
    keep = valid_mask(lengths, x.shape[1])[:, :, None].to(x.dtype)
    pa, pc = (cfg['attn_dropout'], cfg['conv_dropout']) if dropout is not None else (0., 0.)
    a = self_attention(x, lengths, sd, prefix + 'attention.', nb_heads, dropout, pa) * keep
    return conv_feed_forward(a, film, sd, prefix + 'feed_forward.', dropout, pc) * keep


# ----------------------------------------------------------------------------------------------------------------------
# sub-modules
# ----------------------------------------------------------------------------------------------------------------------
def prosody_encoder(sd, hp, frames_energy, frames_pitch, mel_specs, speaker_ids, output_lengths, dropout=None):
    """ProsodyEncoder.forward, model.py:391-464."""
    p = 'prosody_encoder.'
    cfg = hp.prosody_encoder
    D = cfg['hidden_embed_dim']
    T = mel_specs.shape[2]
    dt = mel_specs.dtype
    pos = positional_encoding(output_lengths, T, D, dt)
    energy = conv1d_cl(frames_energy[:, :, None], sd[p + 'energy_embedding.conv.weight'], sd[p + 'energy_embedding.conv.bias'])
    pitch = conv1d_cl(frames_pitch[:, :, None], sd[p + 'pitch_embedding.conv.weight'], sd[p + 'pitch_embedding.conv.bias'])
    x = mel_specs.transpose(1, 2)
    for conv_i, ln_i in ((0, 2), (4, 6), (8, 10)):  # model.py:341-363: conv -> ReLU -> LayerNorm (-> dropout)
        x = conv1d_cl(x, sd[f'{p}convs.{conv_i}.conv.weight'], sd[f'{p}convs.{conv_i}.conv.bias'], relu=True)
        x = F.layer_norm(x, (x.shape[2],), sd[f'{p}convs.{ln_i}.weight'], sd[f'{p}convs.{ln_i}.bias'])
        if dropout is not None:
            x = dropout.rows(x, cfg['conv_dropout'])
    keep = valid_mask(output_lengths, T)[:, :, None].to(dt)
    x = (x + energy + pitch + pos) * keep
    for i in range(cfg['nb_blocks']):
        x = fft_block(x, None, output_lengths, sd, f'{p}blocks.{i}.', cfg['attn_nb_heads'], dropout, cfg)
    pooled = x.sum(dim=1) / output_lengths[:, None]  # model.py:419
    h = pooled + sd[p + 'spk_embedding.weight'][speaker_ids]
    gammas = F.linear(h, sd[p + 'gammas_predictor.linear_layer.weight'], sd[p + 'gammas_predictor.linear_layer.bias'])
    betas = F.linear(h, sd[p + 'betas_predictor.linear_layer.weight'], sd[p + 'betas_predictor.linear_layer.bias'])
    # model.py:322-326,430-461: split per FiLM-ed module, scalar post-multipliers per block
    modules = ((hp.phoneme_encoder['nb_blocks'], hp.phoneme_encoder['hidden_embed_dim']),
               (hp.local_prosody_predictor['nb_blocks'], hp.local_prosody_predictor['conv_channels']),
               (hp.frame_decoder['nb_blocks'], hp.phoneme_encoder['hidden_embed_dim']))
    post = sd.get(p + 'post_multipliers') if hp.post_mult_weight != 0. else None
    films, col, blk = [], 0, 0
    B = h.shape[0]
    for nb, ch in modules:
        g = gammas[:, col:col + nb * ch].reshape(B, nb, ch)
        b = betas[:, col:col + nb * ch].reshape(B, nb, ch)
        if post is not None:
            g = post[0, blk:blk + nb][None, :, None] * g + 1
            b = post[1, blk:blk + nb][None, :, None] * b
        else:
            g = g + 1
        films.append(torch.cat((g, b), dim=2))
        col += nb * ch
        blk += nb
    return pooled, films[0], films[1], films[2]


def speaker_classifier(sd, x):
    """SpeakerClassifier.forward, model.py:276-292 (gradient reversal is the identity in forward, model.py:29-31)."""
    p = 'speaker_classifier.classifier.'
    x = F.relu(F.linear(x, sd[p + '1.linear_layer.weight'], sd[p + '1.linear_layer.bias']))
    x = F.relu(F.linear(x, sd[p + '3.linear_layer.weight'], sd[p + '3.linear_layer.bias']))
    return F.linear(x, sd[p + '5.linear_layer.weight'], sd[p + '5.linear_layer.bias'])


def phoneme_encoder(sd, hp, symbols, film, input_lengths, dropout=None):
    """PhonemeEncoder.forward, model.py:490-509."""
    p = 'phoneme_encoder.'
    cfg = hp.phoneme_encoder
    emb = sd[p + 'symbols_embedding.weight']
    L = symbols.shape[1]
    x = emb[symbols] + positional_encoding(input_lengths, L, cfg['hidden_embed_dim'], emb.dtype)
    x = x * valid_mask(input_lengths, L)[:, :, None].to(emb.dtype)
    for i in range(cfg['nb_blocks']):
        x = fft_block(x, film[:, i, :], input_lengths, sd, f'{p}blocks.{i}.', cfg['attn_nb_heads'], dropout, cfg)
    return x


def local_prosody_predictor(sd, hp, x, film, input_lengths, dropout=None):
    """LocalProsodyPredictor.forward, model.py:549-575."""
    p = 'prosody_predictor.'
    for i in range(hp.local_prosody_predictor['nb_blocks']):
        for conv_i, ln_i in ((0, 2), (4, 6)):
            x = conv1d_cl(x, sd[f'{p}blocks.{i}.{conv_i}.conv.weight'], sd[f'{p}blocks.{i}.{conv_i}.conv.bias'], relu=True)
            x = F.layer_norm(x, (x.shape[2],), sd[f'{p}blocks.{i}.{ln_i}.weight'], sd[f'{p}blocks.{i}.{ln_i}.bias'])
            if dropout is not None:
                x = dropout.rows(x, hp.local_prosody_predictor['conv_dropout'])
        C = x.shape[2]
        assert film.shape[2] == 2 * C
        x = film[:, i, None, :C] * x + film[:, i, None, C:]
    keep = valid_mask(input_lengths, x.shape[1])[:, :, None].to(x.dtype)
    x = x * keep
    y = F.linear(x, sd[p + 'projection.linear_layer.weight'], sd[p + 'projection.linear_layer.bias']) * keep
    return y[:, :, 0], y[:, :, 1], y[:, :, 2]


def gaussian_upsampling(sd, x, durations_float, durations_int, energies, pitch, input_lengths):
    """GaussianUpsamplingModule.forward, model.py:608-662.  The integer contract (bit-exact): cumsum(durations_int),
    T_max = max(cumsum), centres mu_i = d_i/2 + cumsum_{i-1}."""
    p = 'gaussian_upsampling.'
    d = conv1d_cl(durations_float[:, :, None], sd[p + 'duration_projection.conv.weight'], sd[p + 'duration_projection.conv.bias'])
    e = conv1d_cl(energies[:, :, None], sd[p + 'energy_projection.conv.weight'], sd[p + 'energy_projection.conv.bias'])
    f0 = conv1d_cl(pitch[:, :, None], sd[p + 'pitch_projection.conv.weight'], sd[p + 'pitch_projection.conv.bias'])
    x = x + e + f0
    sigma = F.softplus(F.linear(x + d, sd[p + 'projection.0.linear_layer.weight'], sd[p + 'projection.0.linear_layer.bias']))[:, :, 0]
    ok = valid_mask(input_lengths, x.shape[1])
    sigma = torch.where(ok, sigma, torch.ones_like(sigma))
    csum = torch.cumsum(durations_int, dim=1)
    mu = durations_int.to(x.dtype) / 2
    mu[:, 1:] = mu[:, 1:] + csum[:, :-1].to(x.dtype)
    T = int(csum.max())
    t = torch.arange(T, device=x.device, dtype=x.dtype) + 0.5
    # Normal(mu, sigma).log_prob(t) = -(t-mu)^2/(2 sigma^2) - log(sigma) - log(sqrt(2 pi))   (model.py:647-653)
    var = sigma[:, :, None] ** 2
    logp = -((t[None, None, :] - mu[:, :, None]) ** 2) / (2 * var) - sigma[:, :, None].log() - math.log(math.sqrt(2 * math.pi))
    probs = torch.exp(logp) * ok[:, :, None].to(x.dtype)
    weights = probs / (probs.sum(dim=1, keepdim=True) + 1e-20)
    x_up = torch.einsum('blt,bld->btd', weights, x)
    return x_up, weights


def frame_decoder(sd, hp, x, film, output_lengths, dropout=None):
    """FrameDecoder.forward, model.py:689-710."""
    p = 'frame_decoder.'
    cfg = hp.frame_decoder
    T, D = x.shape[1], x.shape[2]
    keep = valid_mask(output_lengths, T)[:, :, None].to(x.dtype)
    x = (x + positional_encoding(output_lengths, T, D, x.dtype)) * keep
    for i in range(cfg['nb_blocks']):
        x = fft_block(x, film[:, i, :], output_lengths, sd, f'{p}blocks.{i}.', cfg['attn_nb_heads'], dropout, cfg)
    mel = F.linear(x, sd[p + 'projection.linear_layer.weight'], sd[p + 'projection.linear_layer.bias']) * keep
    return mel.transpose(1, 2)


# ----------------------------------------------------------------------------------------------------------------------
# top level: forward / loss / inference
# ----------------------------------------------------------------------------------------------------------------------
def forward(sd, hp, inputs, return_intermediates=False, dropout=None):
    """DaftExprt.forward, model.py:755-787 — same 5-tuple.  dropout: None (eval) or a hook with .rows(x, p) / .attn(probs, p)."""
    (symbols, durations_float, durations_int, symbols_energy, symbols_pitch, input_lengths,
     frames_energy, frames_pitch, mel_specs, output_lengths, speaker_ids) = inputs
    prosody_embed, enc_film, pp_film, dec_film = prosody_encoder(sd, hp, frames_energy, frames_pitch, mel_specs, speaker_ids,
                                                                 output_lengths, dropout)
    spk_preds = speaker_classifier(sd, prosody_embed)
    enc = phoneme_encoder(sd, hp, symbols, enc_film, input_lengths, dropout)
    dur, energy, pitch = local_prosody_predictor(sd, hp, enc, pp_film, input_lengths, dropout)
    up, weights = gaussian_upsampling(sd, enc, durations_float, durations_int, symbols_energy, symbols_pitch, input_lengths)
    mel = frame_decoder(sd, hp, up, dec_film, output_lengths, dropout)
    post = sd.get('prosody_encoder.post_multipliers', 1.)
    out = (spk_preds, [post, enc_film, pp_film, dec_film], [dur, energy, pitch, input_lengths], [mel, output_lengths], weights)
    if return_intermediates:
        return out, dict(prosody_embed=prosody_embed, enc_outputs=enc, symbols_upsamp=up)
    return out


def adversarial_weight(iteration, hp):
    """loss.py:22-28."""
    w = iteration * hp.warmup_steps ** -1.5 * hp.adv_max_weight / hp.warmup_steps ** -0.5
    return min(hp.adv_max_weight, w)


def loss(hp, outputs, targets, iteration):
    """DaftExprtLoss.forward, loss.py:30-106 — returns (total, dict of the 7 weighted terms as tensors)."""
    dur_t, energy_t, pitch_t, mel_t, speaker_ids = targets
    spk_preds, film, enc_preds, dec_preds, _ = outputs
    post = film[0]
    dur_p, energy_p, pitch_p, in_len = enc_preds
    mel_p, out_len = dec_preds
    nb = hp.n_mel_channels
    terms = {}
    terms['speaker_loss'] = adversarial_weight(iteration, hp) * F.cross_entropy(spk_preds, speaker_ids)
    if hp.post_mult_weight != 0.:
        terms['post_mult_loss'] = hp.post_mult_weight * torch.linalg.vector_norm(post)
    else:
        terms['post_mult_loss'] = torch.zeros((), dtype=mel_p.dtype)
    terms['duration_loss'] = hp.dur_weight * (((dur_p - dur_t) ** 2).sum(1) / in_len).mean()
    terms['energy_loss'] = hp.energy_weight * (((energy_p - energy_t) ** 2).sum(1) / in_len).mean()
    terms['pitch_loss'] = hp.pitch_weight * (((pitch_p - pitch_t) ** 2).sum(1) / in_len).mean()
    diff = mel_p - mel_t
    terms['mel_spec_l1_loss'] = hp.mel_spec_weight * (diff.abs().sum((1, 2)) / (nb * out_len)).mean()
    terms['mel_spec_l2_loss'] = hp.mel_spec_weight * ((diff ** 2).sum((1, 2)) / (nb * out_len)).mean()
    total = sum(terms.values())
    return total, terms


def duration_to_integer(intervals, hp):
    """extract_features.py:69-111 restated (nb_samples=None branch, as called from model.py:807)."""
    sr, nfft, hop = hp.sampling_rate, hp.filter_length, hp.hop_length
    total = sum((e - b) for b, e in intervals)
    nb_samples = int(total * sr)
    nb_frames = 1 + int((nb_samples - nfft) / hop)
    centres = [int(nfft / 2) + hop * i for i in range(nb_frames)]
    queue = list(intervals)
    out, curr = [], 1
    while curr <= nb_frames:
        b, e = queue.pop(0)  # IndexError if exhausted, like the reference
        if b == e:
            raise ValueError
        bi, ei = int(b * sr), int(e * sr)
        n = len([c for c in centres if bi < c <= ei])
        out.append(n)
        curr += n
    if hp.centered:
        edge = int(nfft / 2 / hop)
        out[0] += edge
        if len(queue) != 0:
            out.append(edge)
        else:
            out[-1] += edge
    return out


def get_int_durations(duration_preds, hp):
    """DaftExprt.get_int_durations, model.py:789-812.  Sequential Python-double accumulation of interval ends."""
    dur_min = (hp.filter_length / hp.sampling_rate) / 2
    duration_preds = torch.where(duration_preds < dur_min, torch.zeros_like(duration_preds), duration_preds)
    out = torch.zeros(duration_preds.shape, dtype=torch.int64)
    for b in range(duration_preds.shape[0]):
        end_prev, idx, intervals = 0., [], []
        for i in range(duration_preds.shape[1]):
            d = duration_preds[b, i].item()
            if d != 0.:
                idx.append(i)
                intervals.append([end_prev, end_prev + d])
                end_prev += d
        ints = duration_to_integer(intervals, hp)
        out[b, idx[:len(ints)]] = torch.tensor(ints, dtype=torch.int64)
    return duration_preds, out


def pitch_shift(pitch, factors, hp, speaker_ids):
    """model.py:814-834."""
    unvoiced = pitch == 0.
    rows = []
    for b in range(pitch.shape[0]):
        st = hp.stats[f'spk {int(speaker_ids[b])}']['pitch']
        hz = torch.exp(st['std'] * pitch[b] + st['mean']) + factors[b]
        rows.append((torch.log(hz) - st['mean']) / st['std'])
    out = torch.stack(rows)
    return torch.where(unvoiced, torch.zeros_like(out), out)


def pitch_multiply(pitch, factors):
    """model.py:836-864."""
    rows = []
    for b in range(pitch.shape[0]):
        voiced = pitch[b] != 0.
        mean = pitch[b][voiced].mean()
        row = pitch[b] + (pitch[b] - mean) * factors[b]
        rows.append(torch.where(voiced, row, torch.zeros_like(row)))
    return torch.stack(rows)


def inference(sd, hp, inputs, pitch_transform):
    """DaftExprt.inference, model.py:866-923."""
    (symbols, dur_factors, energy_factors, pitch_factors, input_lengths,
     energy_refs, pitch_refs, mel_refs, ref_lengths, speaker_ids) = inputs
    _, enc_film, pp_film, dec_film = prosody_encoder(sd, hp, energy_refs, pitch_refs, mel_refs, speaker_ids, ref_lengths)
    enc = phoneme_encoder(sd, hp, symbols, enc_film, input_lengths)
    dur, energy, pitch = local_prosody_predictor(sd, hp, enc, pp_film, input_lengths)
    dur, dur_int = get_int_durations(dur * dur_factors, hp)
    dead = dur_int == 0
    energy = torch.where(dead, torch.zeros_like(energy), energy * energy_factors)
    pitch = torch.where(dead, torch.zeros_like(pitch), pitch)
    if pitch_transform == 'add':
        pitch = pitch_shift(pitch, pitch_factors, hp, speaker_ids)
    elif pitch_transform == 'multiply':
        pitch = pitch_multiply(pitch, pitch_factors)
    else:
        raise NotImplementedError
    up, weights = gaussian_upsampling(sd, enc, dur, dur_int, energy, pitch, input_lengths)
    out_len = dur_int.sum(dim=1)
    assert int(out_len.max()) == up.shape[1]
    mel = frame_decoder(sd, hp, up, dec_film, out_len)
    return [dur, dur_int, energy, pitch, input_lengths], [mel, out_len], weights
