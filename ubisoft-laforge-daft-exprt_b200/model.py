"""Drop-in `DaftExprt` nn.Module over the sm_100a C-ABI (reference `src/daft_exprt/model.py:713-923`).

Same constructor (`DaftExprt(hparams)`), same methods (`parse_batch`, `forward`, `inference`, `get_int_durations`,
`pitch_shift`, `pitch_multiply`), same output tuples and the SAME 193 state-dict keys/shapes as the reference, so existing
checkpoints load unchanged and `train.py` / `generate.py` / `fine_tune.py` / `synthesize.py` run on top of it.

The parameter tree is built from plain torch.nn containers (Conv1d / Linear / LayerNorm / MultiheadAttention / Embedding are
used as parameter holders and initialisers only, in the reference's construction order so that `torch.manual_seed(s)`
yields the same initial weights).  No torch.nn forward is ever called: all math goes through `ops.py` -> libdaftexprt_b200.so.
There is no CPU path; calling forward with CPU tensors raises.
"""
import numpy as np
import torch
from torch import nn

from . import ops

_PE_CACHE = {}


def positional_table(max_len, dim, device):
    """The reference's sinusoid table (model.py:123-130), built with the same CPU torch ops (bit-identical rows), cached
    on the device.  The reference caps it at 5000 rows (IndexError beyond, model.py:147); we grow it on demand."""
    n = max(5000, int(max_len))
    key = (str(device), dim)
    hit = _PE_CACHE.get(key)
    if hit is not None and hit.shape[0] >= n:
        return hit
    pos = torch.arange(0, n, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2).float() * (-np.log(10000.) / dim))
    pe = torch.zeros(n, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    pe = pe.to(device)
    _PE_CACHE[key] = pe
    return pe


# ----------------------------------------------------------------------------------------------------------------------
# parameter holders: names mirror the reference so the state-dict keys are identical
# ----------------------------------------------------------------------------------------------------------------------
def _gain(name):
    return nn.init.calculate_gain(name)


class _Conv(nn.Module):          # ConvNorm1D (model.py:75-94): holder of `.conv`
    def __init__(self, cin, cout, k, gain='linear'):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=k, stride=1, padding=(k - 1) // 2)
        nn.init.xavier_uniform_(self.conv.weight, gain=_gain(gain))


class _Lin(nn.Module):           # LinearNorm (model.py:57-72): holder of `.linear_layer`
    def __init__(self, cin, cout, gain='linear'):
        super().__init__()
        self.linear_layer = nn.Linear(cin, cout)
        nn.init.xavier_uniform_(self.linear_layer.weight, gain=_gain(gain))


class _Slots(nn.Module):
    """Children registered under the numeric names an nn.Sequential would give them (ReLU/Dropout slots stay empty)."""

    def __init__(self, **named):
        super().__init__()
        for k, v in named.items():
            self.add_module(k.lstrip('_'), v)

    def __getitem__(self, i):
        return self._modules[str(i)]


class _Attention(nn.Module):     # MultiHeadAttention (model.py:153-193)
    def __init__(self, dim, heads, p):
        super().__init__()
        self.multi_head_attention = nn.MultiheadAttention(dim, heads, p)
        self.layer_norm = nn.LayerNorm(dim)
        self.nb_heads, self.p = heads, p

    def forward(self, x, lens):
        m = self.multi_head_attention
        return ops.AttentionSubLayer.apply(x, lens, m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias,
                                           self.layer_norm.weight, self.layer_norm.bias, self.nb_heads,
                                           self.p if self.training else 0.0)


class _ConvFF(nn.Module):        # PositionWiseConvFF (model.py:196-237)
    def __init__(self, dim, channels, k, p):
        super().__init__()
        self.convs = _Slots(_0=_Conv(dim, channels, k, 'relu'), _2=_Conv(channels, dim, k, 'linear'))
        self.layer_norm = nn.LayerNorm(dim)
        self.p = p

    def forward(self, x, film, lens):
        c1, c2 = self.convs[0].conv, self.convs[2].conv
        return ops.ConvFFSubLayer.apply(x, lens, c1.weight, c1.bias, c2.weight, c2.bias, self.layer_norm.weight,
                                        self.layer_norm.bias, film, self.p if self.training else 0.0)


class FFTBlock(nn.Module):       # model.py:240-264
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg['hidden_embed_dim'], cfg['attn_nb_heads'], cfg['attn_dropout'])
        self.feed_forward = _ConvFF(cfg['hidden_embed_dim'], cfg['conv_channels'], cfg['conv_kernel'], cfg['conv_dropout'])

    def forward(self, x, film, lens):
        return self.feed_forward(self.attention(x, lens), film, lens)


class ProsodyEncoder(nn.Module):  # model.py:295-464
    def __init__(self, hp):
        super().__init__()
        cfg = hp.prosody_encoder
        D, C, k, p = cfg['hidden_embed_dim'], cfg['conv_channels'], cfg['conv_kernel'], cfg['conv_dropout']
        self.post_mult_weight = hp.post_mult_weight
        self.film_layout = ((hp.phoneme_encoder['nb_blocks'], hp.phoneme_encoder['hidden_embed_dim']),
                            (hp.local_prosody_predictor['nb_blocks'], hp.local_prosody_predictor['conv_channels']),
                            (hp.frame_decoder['nb_blocks'], hp.phoneme_encoder['hidden_embed_dim']))
        self.dim, self.p = D, p
        self.energy_embedding = _Conv(1, D, k)
        self.pitch_embedding = _Conv(1, D, k)
        self.convs = _Slots(_0=_Conv(hp.n_mel_channels, C, k, 'relu'), _2=nn.LayerNorm(C),
                            _4=_Conv(C, C, k, 'relu'), _6=nn.LayerNorm(C),
                            _8=_Conv(C, D, k, 'relu'), _10=nn.LayerNorm(D))
        self.blocks = nn.ModuleList([FFTBlock(cfg) for _ in range(cfg['nb_blocks'])])
        self.spk_embedding = nn.Embedding(hp.n_speakers, D)
        nn.init.xavier_uniform_(self.spk_embedding.weight.data)
        nf = sum(nb * ch for nb, ch in self.film_layout)
        self.gammas_predictor = _Lin(D, nf)
        self.betas_predictor = _Lin(D, nf)
        if self.post_mult_weight != 0.:
            self.post_multipliers = nn.Parameter(torch.empty(2, sum(nb for nb, _ in self.film_layout)))
            nn.init.xavier_uniform_(self.post_multipliers, gain=_gain('linear'))
        else:
            self.post_multipliers = 1.

    def forward(self, frames_energy, frames_pitch, mel_specs, speaker_ids, output_lengths):
        T = mel_specs.shape[2]
        p = self.p if self.training else 0.0
        c = self.convs
        x = ops.PreNet.apply(mel_specs, output_lengths, c[0].conv.weight, c[0].conv.bias, c[2].weight, c[2].bias,
                             c[4].conv.weight, c[4].conv.bias, c[6].weight, c[6].bias,
                             c[8].conv.weight, c[8].conv.bias, c[10].weight, c[10].bias, p)
        pe = positional_table(T, self.dim, mel_specs.device)
        x = ops.FrameInput.apply(x, output_lengths, pe, frames_energy, frames_pitch,
                                 self.energy_embedding.conv.weight, self.energy_embedding.conv.bias,
                                 self.pitch_embedding.conv.weight, self.pitch_embedding.conv.bias)
        for block in self.blocks:
            x = block(x, None, output_lengths)
        pooled, h = ops.MeanPoolSpeaker.apply(x, output_lengths, speaker_ids, self.spk_embedding.weight)
        post = self.post_multipliers if self.post_mult_weight != 0. else None
        nbs, chs = [nb for nb, _ in self.film_layout], [ch for _, ch in self.film_layout]
        films = ops.FilmHead.apply(h, self.gammas_predictor.linear_layer.weight, self.gammas_predictor.linear_layer.bias,
                                   self.betas_predictor.linear_layer.weight, self.betas_predictor.linear_layer.bias,
                                   post, nbs, chs)
        per_module, k = [], 0
        for nb in nbs:
            per_module.append(list(films[k:k + nb]))
            k += nb
        return pooled, per_module


class SpeakerClassifier(nn.Module):  # model.py:267-292 (+ gradient reversal, model.py:27-54)
    def __init__(self, hp):
        super().__init__()
        D = hp.prosody_encoder['hidden_embed_dim']
        self.lambda_ = hp.lambda_reversal
        self.classifier = _Slots(_1=_Lin(D, D, 'relu'), _3=_Lin(D, D, 'relu'), _5=_Lin(D, hp.n_speakers - 1, 'linear'))

    def forward(self, x):
        l1, l3, l5 = (self.classifier[i].linear_layer for i in (1, 3, 5))
        x = ops.Linear.apply(x, l1.weight, l1.bias, True, -self.lambda_)   # reversal: dx = -lambda * grad
        x = ops.Linear.apply(x, l3.weight, l3.bias, True, 1.0)
        return ops.Linear.apply(x, l5.weight, l5.bias, False, 1.0)


class PhonemeEncoder(nn.Module):  # model.py:467-509
    def __init__(self, hp):
        super().__init__()
        cfg = hp.phoneme_encoder
        self.dim = cfg['hidden_embed_dim']
        self.symbols_embedding = nn.Embedding(hp.n_symbols, self.dim)
        nn.init.xavier_uniform_(self.symbols_embedding.weight.data)
        self.blocks = nn.ModuleList([FFTBlock(cfg) for _ in range(cfg['nb_blocks'])])

    def forward(self, symbols, films, input_lengths):
        pe = positional_table(symbols.shape[1], self.dim, symbols.device)
        x = ops.EmbedPE.apply(symbols, input_lengths, self.symbols_embedding.weight, pe)
        for i, block in enumerate(self.blocks):
            x = block(x, films[i], input_lengths)
        return x


class LocalProsodyPredictor(nn.Module):  # model.py:512-575
    def __init__(self, hp):
        super().__init__()
        D = hp.phoneme_encoder['hidden_embed_dim']
        cfg = hp.local_prosody_predictor
        C, k = cfg['conv_channels'], cfg['conv_kernel']
        self.p = cfg['conv_dropout']
        self.blocks = nn.ModuleList([_Slots(_0=_Conv(D if i == 0 else C, C, k, 'relu'), _2=nn.LayerNorm(C),
                                            _4=_Conv(C, C, k, 'relu'), _6=nn.LayerNorm(C)) for i in range(cfg['nb_blocks'])])
        self.projection = _Lin(C, 3, 'linear')

    def forward(self, x, films, input_lengths):
        p = self.p if self.training else 0.0
        for i, b in enumerate(self.blocks):
            last = i == len(self.blocks) - 1   # the reference masks once, after the last block's FiLM (model.py:567-568)
            x = ops.PredictorBlock.apply(x, input_lengths if last else None, films[i], b[0].conv.weight, b[0].conv.bias, b[2].weight,
                                         b[2].bias, b[4].conv.weight, b[4].conv.bias, b[6].weight, b[6].bias, p)
        out = ops.PredictorHead.apply(x, input_lengths, self.projection.linear_layer.weight, self.projection.linear_layer.bias)
        return out[0], out[1], out[2]


class GaussianUpsamplingModule(nn.Module):  # model.py:578-662
    def __init__(self, hp):
        super().__init__()
        D = hp.phoneme_encoder['hidden_embed_dim']
        k = hp.gaussian_upsampling_module['conv_kernel']
        self.duration_projection = _Conv(1, D, k)
        self.energy_projection = _Conv(1, D, k)
        self.pitch_projection = _Conv(1, D, k)
        self.projection = _Slots(_0=_Lin(D, 1, 'relu'))

    def forward(self, x, durations_float, durations_int, energies, pitch, input_lengths, nb_frames_max=None):
        d, e, f = self.duration_projection.conv, self.energy_projection.conv, self.pitch_projection.conv
        r = self.projection[0].linear_layer
        up, weights, _csum, totals = ops.GaussUpsample.apply(x, durations_float, durations_int, energies, pitch, input_lengths,
                                                             d.weight, d.bias, e.weight, e.bias, f.weight, f.bias,
                                                             r.weight, r.bias, nb_frames_max)
        return up, weights, totals


class FrameDecoder(nn.Module):  # model.py:665-710
    def __init__(self, hp):
        super().__init__()
        D = hp.phoneme_encoder['hidden_embed_dim']
        hp.frame_decoder['hidden_embed_dim'] = D   # the reference mutates hparams the same way (model.py:675)
        cfg = hp.frame_decoder
        self.dim = D
        self.blocks = nn.ModuleList([FFTBlock(cfg) for _ in range(cfg['nb_blocks'])])
        self.projection = _Lin(D, hp.n_mel_channels, 'linear')

    def forward(self, x, films, output_lengths):
        pe = positional_table(x.shape[1], self.dim, x.device)
        x = ops.FrameInput.apply(x, output_lengths, pe, None, None, None, None, None, None)
        for i, block in enumerate(self.blocks):
            x = block(x, films[i], output_lengths)
        return ops.MelProjection.apply(x, output_lengths, self.projection.linear_layer.weight, self.projection.linear_layer.bias)


class DaftExprt(nn.Module):
    """Reference `DaftExprt` (model.py:713-923) on hand-written sm_100a kernels."""

    def __init__(self, hparams):
        super().__init__()
        self.prosody_encoder = ProsodyEncoder(hparams)
        self.speaker_classifier = SpeakerClassifier(hparams)
        self.phoneme_encoder = PhonemeEncoder(hparams)
        self.prosody_predictor = LocalProsodyPredictor(hparams)
        self.gaussian_upsampling = GaussianUpsamplingModule(hparams)
        self.frame_decoder = FrameDecoder(hparams)
        self._stats_cache = None
        # weights written behind autograd's back (checkpoint load into re-pointed storage) must not meet stale packs
        self.register_load_state_dict_post_hook(lambda module, incompatible: ops.invalidate_packed_weights())

    # -- model.py:727-753 ------------------------------------------------------------------------------------------------
    def parse_batch(self, gpu, batch):
        dev = torch.device('cuda', gpu) if isinstance(gpu, int) else torch.device(gpu)
        if hasattr(batch, 'to_device'):
            # data.FlatBatch: the 11 tensors live in ONE (pinned) byte buffer -> ONE host->device copy, device views, no casts
            self._check_ids(batch.tensors()[0], batch.tensors()[10], training_batch=True)
            inputs = batch.to_device(dev)
            targets = (inputs[1], inputs[3], inputs[4], inputs[8], inputs[10])
            return inputs, targets, (batch.feature_dirs, batch.feature_files)
        else:
            (symbols, durations_float, durations_int, symbols_energy, symbols_pitch, input_lengths, frames_energy, frames_pitch,
             mel_specs, output_lengths, speaker_ids, feature_dirs, feature_files) = batch
            self._check_ids(symbols, speaker_ids, training_batch=True)
            f = lambda t: t.to(dev, non_blocking=True).float()
            i = lambda t: t.to(dev, non_blocking=True).long()
            symbols, durations_int, input_lengths = i(symbols), i(durations_int), i(input_lengths)
            output_lengths, speaker_ids = i(output_lengths), i(speaker_ids)
            durations_float, symbols_energy, symbols_pitch = f(durations_float), f(symbols_energy), f(symbols_pitch)
            frames_energy, frames_pitch, mel_specs = f(frames_energy), f(frames_pitch), f(mel_specs)
        inputs = (symbols, durations_float, durations_int, symbols_energy, symbols_pitch, input_lengths,
                  frames_energy, frames_pitch, mel_specs, output_lengths, speaker_ids)
        targets = (durations_float, symbols_energy, symbols_pitch, mel_specs, speaker_ids)
        return inputs, targets, (feature_dirs, feature_files)

    def _check_ids(self, symbols, speaker_ids, training_batch):
        """The reference raises on out-of-range ids (nn.Embedding IndexError, model.py:424,497; cross_entropy 'Target out of
        bounds', loss.py:60 — the classifier has n_speakers - 1 classes, model.py:274).  The kernels clamp instead of faulting, so
        the range is checked here, on the HOST copy of the batch (no device sync); device-resident ids are checked by
        `inference` through its single read-back."""
        if symbols.device.type != 'cpu' or symbols.numel() == 0:
            return
        n_sym = self.phoneme_encoder.symbols_embedding.num_embeddings
        n_spk = self.prosody_encoder.spk_embedding.num_embeddings
        lo, hi = int(symbols.min()), int(symbols.max())
        if lo < 0 or hi >= n_sym:
            raise IndexError(f'symbol id out of range [0, {n_sym}): min {lo}, max {hi}')
        lo, hi = int(speaker_ids.min()), int(speaker_ids.max())
        lim = n_spk - 1 if training_batch else n_spk
        if lo < 0 or hi >= lim:
            raise IndexError(f'speaker id out of range [0, {lim}): min {lo}, max {hi}'
                             + (' (the speaker classifier has n_speakers - 1 classes)' if training_batch else ''))

    @staticmethod
    def _stack_films(films):
        with torch.no_grad():
            return torch.stack([f.detach() for f in films], dim=1)

    # -- model.py:755-787 ------------------------------------------------------------------------------------------------
    def _gemm_weights(self):
        ws = self.__dict__.get('_gemm_weight_list')
        if ws is None:
            ws = [p for p in self.parameters() if p.dim() in (2, 3)]
            self.__dict__['_gemm_weight_list'] = ws
        return ws

    def forward(self, inputs):
        (symbols, durations_float, durations_int, symbols_energy, symbols_pitch, input_lengths, frames_energy, frames_pitch,
         mel_specs, output_lengths, speaker_ids) = inputs
        input_lengths, output_lengths = input_lengths.detach().contiguous(), output_lengths.detach().contiguous()
        ops.prepack(self._gemm_weights())   # one launch refreshes every weight pack the optimiser step made stale
        prosody_embed, (enc_film, pp_film, dec_film) = self.prosody_encoder(frames_energy, frames_pitch, mel_specs,
                                                                             speaker_ids.contiguous(), output_lengths)
        spk_preds = self.speaker_classifier(prosody_embed)
        enc_outputs = self.phoneme_encoder(symbols, enc_film, input_lengths)
        duration_preds, energy_preds, pitch_preds = self.prosody_predictor(enc_outputs, pp_film, input_lengths)
        # training: T_max is the reference mel length (sum(durations_int) == output_lengths, data_loader.py:128) -> no sync
        symbols_upsamp, weights, _ = self.gaussian_upsampling(enc_outputs, durations_float, durations_int, symbols_energy,
                                                              symbols_pitch, input_lengths, mel_specs.shape[2])
        mel_spec_preds = self.frame_decoder(symbols_upsamp, dec_film, output_lengths)
        film_params = [self.prosody_encoder.post_multipliers, self._stack_films(enc_film), self._stack_films(pp_film),
                       self._stack_films(dec_film)]
        encoder_preds = [duration_preds, energy_preds, pitch_preds, input_lengths]
        decoder_preds = [mel_spec_preds, output_lengths]
        return spk_preds, film_params, encoder_preds, decoder_preds, weights

    # -- model.py:789-812 + extract_features.py:69-111 ---------------------------------------------------------------------
    def get_int_durations(self, duration_preds, hparams, dur_factors=None):
        d = ops._check_input(duration_preds)
        B, L = d.shape
        out = torch.empty_like(d)
        dint = torch.empty(B, L, device=d.device, dtype=torch.int64)
        totals = torch.empty(B, device=d.device, dtype=torch.int64)
        err = torch.empty(B, device=d.device, dtype=torch.int32)
        ops._call('dx_int_durations', ops._p(d), ops._p(dur_factors), None, ops._p(out), ops._p(dint), ops._p(totals),
                  ops._p(err), B, L, int(hparams.sampling_rate), int(hparams.filter_length), int(hparams.hop_length),
                  int(bool(hparams.centered)), ops._st())
        self._last_int_dur_status = (totals, err)
        return out, dint

    def _pitch_stats(self, hparams, device):
        ids = sorted(int(k.split(' ')[1]) for k in hparams.stats if k.startswith('spk '))
        n = (max(ids) + 1) if ids else 1
        table = torch.zeros(n, 2, dtype=torch.float32)
        for i in ids:
            table[i, 0] = hparams.stats[f'spk {i}']['pitch']['mean']
            table[i, 1] = hparams.stats[f'spk {i}']['pitch']['std']
        return table.to(device)

    # -- model.py:814-834 ------------------------------------------------------------------------------------------------
    def pitch_shift(self, pitch_preds, pitch_factors, hparams, speaker_ids):
        B, L = pitch_preds.shape
        stats = self._pitch_stats(hparams, pitch_preds.device)
        ops._call('dx_pitch_shift', ops._p(pitch_preds), ops._p(ops._check_input(pitch_factors)), ops._p(speaker_ids),
                  ops._p(stats), B, L, ops._st())
        return pitch_preds

    # -- model.py:836-864 ------------------------------------------------------------------------------------------------
    def pitch_multiply(self, pitch_preds, pitch_factors):
        B, L = pitch_preds.shape
        ops._call('dx_pitch_multiply', ops._p(pitch_preds), ops._p(ops._check_input(pitch_factors)), B, L, ops._st())
        return pitch_preds

    # -- model.py:866-923 ------------------------------------------------------------------------------------------------
    def inference(self, inputs, pitch_transform, hparams):
        (symbols, dur_factors, energy_factors, pitch_factors, input_lengths, energy_refs, pitch_refs, mel_spec_refs, ref_lengths,
         speaker_ids) = inputs
        if pitch_transform not in ('add', 'multiply'):
            raise NotImplementedError
        ops.prepack(self._gemm_weights())
        input_lengths, ref_lengths = input_lengths.contiguous(), ref_lengths.contiguous()
        speaker_ids = speaker_ids.contiguous()
        _, (enc_film, pp_film, dec_film) = self.prosody_encoder(energy_refs, pitch_refs, mel_spec_refs, speaker_ids, ref_lengths)
        enc_outputs = self.phoneme_encoder(symbols, enc_film, input_lengths)
        duration_preds, energy_preds, pitch_preds = self.prosody_predictor(enc_outputs, pp_film, input_lengths)
        B, L = duration_preds.shape
        duration_preds, durations_int = self.get_int_durations(duration_preds.detach(), hparams, ops._check_input(dur_factors))
        energy_preds = energy_preds.detach().clone()
        pitch_preds = pitch_preds.detach().clone()
        ops._call('dx_inference_adjust', ops._p(energy_preds), ops._p(pitch_preds), ops._p(ops._check_input(energy_factors)),
                  ops._p(durations_int), B, L, ops._st())
        if pitch_transform == 'add':
            pitch_preds = self.pitch_shift(pitch_preds, pitch_factors, hparams, speaker_ids)
        else:
            pitch_preds = self.pitch_multiply(pitch_preds, pitch_factors)
        totals, err = self._last_int_dur_status
        # ONE read-back: T_max, the error flag and the id ranges (the kernels clamp ids; the reference raises IndexError)
        status = torch.stack((totals.max(), err.max().long(), symbols.max(), symbols.min(), speaker_ids.max(), speaker_ids.min())).tolist()
        n_sym = self.phoneme_encoder.symbols_embedding.num_embeddings
        n_spk = self.prosody_encoder.spk_embedding.num_embeddings
        if status[3] < 0 or status[2] >= n_sym or status[5] < 0 or status[4] >= n_spk:
            raise IndexError(f'symbol / speaker id out of range: symbols [{status[3]}, {status[2]}] vs {n_sym}, '
                             f'speakers [{status[5]}, {status[4]}] vs {n_spk}')
        if status[1] != 0:
            raise IndexError('get_int_durations: predicted durations too short to cover the frame grid '
                             '(the reference raises here too, extract_features.py:88-92)')
        nb_frames_max = int(status[0])
        symbols_upsamp, weights, output_lengths = self.gaussian_upsampling(enc_outputs, duration_preds, durations_int,
                                                                           energy_preds, pitch_preds, input_lengths, nb_frames_max)
        assert nb_frames_max == symbols_upsamp.size(1)   # model.py:914
        mel_spec_preds = self.frame_decoder(symbols_upsamp, dec_film, output_lengths)
        encoder_preds = [duration_preds, durations_int, energy_preds, pitch_preds, input_lengths]
        decoder_preds = [mel_spec_preds, output_lengths]
        return encoder_preds, decoder_preds, weights


def reference_state_shapes(n_speakers=12, n_symbols=76, n_mel_channels=80, hparams=None):
    """{state-dict key: shape} of the module (== the reference's 193 tensors); built on the meta device, no GPU needed."""
    from .hparams import default_hparams
    hp = hparams or default_hparams(n_speakers=n_speakers, n_symbols=n_symbols, n_mel_channels=n_mel_channels)
    with torch.device('meta'):
        m = DaftExprt(hp)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}
