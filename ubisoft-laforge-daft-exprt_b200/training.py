"""Callers of the hot path, in the reference's own loop shape.

`train_epoch` is the body of the reference's hot loop (`train.py:368-401,486-491`): parse_batch, forward, loss / accumulation_steps,
backward, and every `accumulation_steps` micro-batches clip_grad_norm_ + optimizer.step + zero_grad + learning-rate update.  It runs
either eagerly (any torch optimiser, or `FlatAdam` + `FlatGradSync`) or through `graph.GraphedTrainStep` (same arithmetic, one
CUDA-graph replay per micro-batch).  Hygiene differences vs the reference, none of which change the math:
  * one gradient all-reduce per OPTIMISER step instead of one per micro-batch (the reference's DDP has no `no_sync()`,
    train.py:293,391);
  * the loss terms stay on the device; ONE read-back per optimiser step (the reference does 8 `.item()` per micro-batch,
    loss.py:102-104, train.py:382).

`validate_sharded` is `validate` (`train.py:193-233`) with the validation set SHARDED over ranks (the reference makes every rank
run the full set, `data_loader.py:240`): rank r takes batches r, r + world, ..., the six loss sums and the batch count are
combined with ONE all-reduce, and the means returned are those of the full set on every rank.
"""
import math

import torch
import torch.distributed as dist

from .graph import reference_lr_schedule
from .loss import TERMS

VAL_TERMS = ('duration_loss', 'energy_loss', 'pitch_loss', 'mel_spec_l1_loss', 'mel_spec_l2_loss')   # train.py:205-208


def _unwrap(model):
    return model.module if hasattr(model, 'module') else model


def validate_sharded(gpu, model, criterion, val_loader, hparams=None, group=None, keep_outputs=True):
    """-> (val_loss, val_indiv_loss, val_targets, val_outputs) like train.py:193-233; the targets / outputs lists hold this
    rank's shard only (the reference uses them for TensorBoard plots on rank 0)."""
    m = _unwrap(model)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    was_training = m.training
    m.eval()
    acc = None
    val_targets, val_outputs = [], []
    with torch.no_grad():
        for i, batch in enumerate(val_loader):
            if i % world != rank:
                continue
            inputs, targets, _ = m.parse_batch(gpu, batch)
            outputs = model(inputs)
            out = criterion.forward_device(outputs, targets, 0)      # float32[8] on the device, no sync
            row = torch.cat((out.detach().double(), torch.ones(1, device=out.device, dtype=torch.float64)))
            acc = row if acc is None else acc + row
            if keep_outputs:
                val_targets.append(targets)
                val_outputs.append(outputs)
    if acc is None:   # this rank got no batch (fewer batches than ranks)
        dev = torch.device('cuda', gpu) if isinstance(gpu, int) else torch.device(gpu)
        acc = torch.zeros(9, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)     # the ONE collective of the validation pass
    vals = acc.tolist()                                              # the ONE read-back
    n = max(vals[8], 1.0)
    val_loss = vals[7] / n
    val_indiv_loss = {k: vals[TERMS.index(k)] / n for k in VAL_TERMS}
    m.train(was_training)
    return val_loss, val_indiv_loss, val_targets, val_outputs


def train_epoch(gpu, model, criterion, optimizer, batches, hparams, iteration, sync=None, graphed=None, on_step=None):
    """The reference's inner loop (train.py:368-401,486-491) over `batches`; returns the iteration reached.

    eager mode : `optimizer` is any torch.optim-style object (torch Adam on the module, or FlatAdam); `sync` (FlatGradSync,
                 optional) provides the flat bucket + the single all-reduce per optimiser step.
    graph mode : `graphed` is a GraphedTrainStep built with accumulation_steps = hparams.accumulation_steps.
    `on_step(iteration, tot_loss, indiv_loss, grad_norm, learning_rate)` is called once per optimiser step (the reference logs here)."""
    m = _unwrap(model)
    acc_steps = int(getattr(hparams, 'accumulation_steps', 1))
    lr_of = reference_lr_schedule(hparams)
    clip = float(getattr(hparams, 'grad_clip_thresh', float('inf')))
    micro, pending = 0, []
    if graphed is None:
        optimizer.zero_grad()
    for batch in batches:
        inputs, targets, _ = m.parse_batch(gpu, batch)
        if graphed is not None:
            pending.append(graphed.step(inputs, targets, iteration).clone())
            stepped = graphed.stepped
            grad_norm = None
        else:
            out = criterion.forward_device(model(inputs), targets, iteration)
            (out[7] / acc_steps).backward()                          # train.py:379,391
            pending.append(out.detach())
            micro += 1
            stepped = micro == acc_steps
            grad_norm = None
            if stepped:
                micro = 0
                if sync is not None:
                    sync.all_reduce_mean()
                if hasattr(optimizer, 'grad_clip_thresh'):            # FlatAdam: clip on the device inside step()
                    optimizer.grad_clip_thresh = clip
                    optimizer.step()
                else:
                    grad_norm = torch.nn.utils.clip_grad_norm_(m.parameters(), clip)   # train.py:399
                    optimizer.step()
                optimizer.zero_grad()
        if stepped:
            if on_step is not None:
                terms = (torch.stack(pending).sum(0) / acc_steps).tolist()           # ONE read-back per optimiser step
                lr = next(g['lr'] for g in optimizer.param_groups if g['lr'] is not None)
                if grad_norm is None and hasattr(optimizer, 'wants_clip') and optimizer.wants_clip():
                    grad_norm = optimizer.last_grad_norm()
                if not math.isnan(terms[7]):
                    on_step(iteration, terms[7], {k: terms[i] for i, k in enumerate(TERMS)}, grad_norm, lr)
            pending = []
            iteration += 1                                            # train.py:475
            new_lr = lr_of(iteration)                                 # train.py:491-494
            for g in optimizer.param_groups:
                g['lr'] = new_lr
    return iteration
