"""Drop-in `DaftExprtLoss` (reference `src/daft_exprt/loss.py:6-106`) as one fused reduction on the device.

Same constructor `(gpu, hparams)`, same call `(outputs, targets, iteration)`, same return `(loss, individual_loss dict)`.
The seven weighted terms and the total land in ONE 8-float device buffer; the dict of Python floats the reference returns
(loss.py:102-104, seven `.item()` syncs) is produced with a single read-back.  `forward_device` returns the buffer
without any sync for callers that do not need the floats every step.
"""
import torch
from torch import nn

from . import ops

TERMS = ('speaker_loss', 'post_mult_loss', 'duration_loss', 'energy_loss', 'pitch_loss', 'mel_spec_l1_loss', 'mel_spec_l2_loss')


class DaftExprtLoss(nn.Module):
    def __init__(self, gpu, hparams):
        super().__init__()
        self.nb_channels = hparams.n_mel_channels
        self.warmup_steps = hparams.warmup_steps
        self.adv_max_weight = hparams.adv_max_weight
        self.post_mult_weight = hparams.post_mult_weight
        self.dur_weight = hparams.dur_weight
        self.energy_weight = hparams.energy_weight
        self.pitch_weight = hparams.pitch_weight
        self.mel_spec_weight = hparams.mel_spec_weight

    def update_adversarial_weight(self, iteration):
        # loss.py:22-28
        weight_iter = iteration * self.warmup_steps ** -1.5 * self.adv_max_weight / self.warmup_steps ** -0.5
        return min(self.adv_max_weight, weight_iter)

    def forward_device(self, outputs, targets, iteration):
        """-> float32[8] device tensor {7 weighted terms, total}; differentiable through element 7."""
        duration_targets, energy_targets, pitch_targets, mel_spec_targets, speaker_ids = targets
        speaker_preds, film_params, encoder_preds, decoder_preds, _ = outputs
        post_multipliers = film_params[0]
        duration_preds, energy_preds, pitch_preds, input_lengths = encoder_preds
        mel_spec_preds, output_lengths = decoder_preds
        post = post_multipliers if (self.post_mult_weight != 0. and torch.is_tensor(post_multipliers)) else None
        weights = (self.update_adversarial_weight(iteration), self.post_mult_weight, self.dur_weight, self.energy_weight,
                   self.pitch_weight, self.mel_spec_weight)
        return ops.Loss.apply(speaker_preds, post, duration_preds, energy_preds, pitch_preds, mel_spec_preds,
                              speaker_ids.contiguous(), duration_targets, energy_targets, pitch_targets, mel_spec_targets,
                              input_lengths.contiguous(), output_lengths.contiguous(), weights)

    def forward(self, outputs, targets, iteration):
        out = self.forward_device(outputs, targets, iteration)
        vals = out.detach().tolist()   # the single device->host read of the step
        individual_loss = {k: vals[i] for i, k in enumerate(TERMS)}
        return out[7], individual_loss


class LossReadback:
    """One-step-delayed read-back of the 8 loss floats: the device->host copy of step i is enqueued right behind step i (pinned
    ring slot, async) and its VALUES are collected when step i + 1 is submitted, so the host never stalls on the step it has just
    launched (the reference blocks on 8 `.item()` calls per micro-batch, loss.py:102-104, train.py:382).  `flush()` returns what is
    still in flight (call it before reading the clock / at the end of an epoch)."""

    def __init__(self, slots=4):
        self.host = torch.zeros(slots, 8, dtype=torch.float32).pin_memory()
        self.pending = []          # (slot, event)
        self.k = 0

    def submit(self, out8):
        """Enqueue the copy of this step's loss tensor; returns the floats of the PREVIOUS submitted step (None for the first)."""
        slot = self.k % self.host.shape[0]
        self.k += 1
        self.host[slot].copy_(out8.detach(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        prev = self._collect() if self.pending else None
        self.pending.append((slot, ev))
        return prev

    def _collect(self):
        slot, ev = self.pending.pop(0)
        ev.synchronize()
        return self.host[slot].tolist()

    def flush(self):
        out = []
        while self.pending:
            out.append(self._collect())
        return out
