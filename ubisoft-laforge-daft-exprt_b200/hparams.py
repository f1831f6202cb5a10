"""The slice of the reference's `HyperParams` object (hparams.py:19-244) that the hot path reads.

`DaftExprt` / `DaftExprtLoss` accept the reference's own `HyperParams` instance unchanged; this light namespace carries the
same attribute names and default values for use without the reference's on-disk requirements (MFA model files,
hparams.py:219-230).  Values are the reference defaults (hparams.py:36-128)."""
import copy
from types import SimpleNamespace

_DEFAULTS = dict(
    # features (hparams.py:40-46)
    centered=True, sampling_rate=22050, n_mel_channels=80, filter_length=1024, hop_length=256,
    # training (hparams.py:58-67)
    seed=1234, batch_size=16, accumulation_steps=3, grad_clip_thresh=float('inf'),
    # loss weights (hparams.py:70-76)
    lambda_reversal=1., adv_max_weight=1e-2, post_mult_weight=1e-3, dur_weight=1., energy_weight=1., pitch_weight=1.,
    mel_spec_weight=1.,
    # optimiser (hparams.py:79-87)
    betas=(0.9, 0.98), epsilon=1e-9, weight_decay=1e-6, initial_learning_rate=1e-4, max_learning_rate=1e-3, warmup_steps=10000,
    # modules (hparams.py:90-128)
    prosody_encoder=dict(nb_blocks=4, hidden_embed_dim=128, attn_nb_heads=8, attn_dropout=0.1, conv_kernel=3,
                         conv_channels=1024, conv_dropout=0.1),
    phoneme_encoder=dict(nb_blocks=4, hidden_embed_dim=128, attn_nb_heads=2, attn_dropout=0.1, conv_kernel=3,
                         conv_channels=1024, conv_dropout=0.1),
    local_prosody_predictor=dict(nb_blocks=1, conv_kernel=3, conv_channels=256, conv_dropout=0.1),
    gaussian_upsampling_module=dict(conv_kernel=3),
    frame_decoder=dict(nb_blocks=4, attn_nb_heads=2, attn_dropout=0.1, conv_kernel=3, conv_channels=1024, conv_dropout=0.1),
)


def default_hparams(n_speakers=12, n_symbols=76, stats=None, **overrides):
    """n_speakers follows the reference's convention: number of speaker ids + 1 (hparams.py:200-201)."""
    d = copy.deepcopy(_DEFAULTS)
    d.update(n_speakers=n_speakers, n_symbols=n_symbols, stats=stats or {})
    d.update(overrides)
    return SimpleNamespace(**d)
