"""Host -> device hand-over of collated batches: bucketed padding, ONE pinned buffer, ONE H2D copy.

The reference's loop (`train.py:368-378`) collates every batch to ITS OWN max lengths (`data_loader.py:140-211`), then calls
`model.parse_batch(gpu, batch)` — eleven `.cuda(non_blocking=True)` copies — and the model, back to back on one stream.  Three
things are provided here, each usable on its own with the reference's `DataLoader`:

  * `BucketedCollate(base_collate, ...)` wraps the reference's `DaftExprtDataCollate` and right-pads its output up to a small grid
    of (L, T) bucket shapes, so that a handful of captured CUDA graphs (`graph.GraphedTrainStep`) covers the whole length
    distribution instead of one capture per batch.  The padded rows cost little: every kernel takes the true lengths and skips
    dead tiles.  Semantics: the model's result is exactly what the reference computes ON THE SAME PADDED TENSORS (i.e. as if a
    longer utterance had been in the batch); relative to the unpadded batch only the last <= 3 frames of the single longest
    utterance can move — the batch-composition dependence every other utterance of a reference batch already has
    (the conv stacks run unmasked over the zero-padded layout).
  * `FlatBatch`: the 11 tensors of a batch as views of ONE contiguous byte buffer.  `DataLoader(pin_memory=True)` pins it with one
    call (`FlatBatch.pin_memory`), `DaftExprt.parse_batch` moves it with ONE cudaMemcpyAsync and re-creates the tensors as
    device views (no dtype-conversion kernels: the collate already produces int64 / float32).  It unpacks like the reference's
    13-tuple, so `model.parse_batch(gpu, batch)` and `fine_tune.py` keep working.
  * `BatchPrefetcher` issues `parse_batch` for the NEXT batch on a side stream while the current step computes.

    loader = DataLoader(train_set, ..., pin_memory=True, collate_fn=BucketedCollate(DaftExprtDataCollate(hparams)))
    pre = BatchPrefetcher(model, gpu)
    pre.submit(next(it))
    for batch in it:
        inputs, targets, file_ids = pre.get()      # waits (on the device) for the copy of THIS step
        pre.submit(batch)                          # next step's copy overlaps this step's kernels
        ...

`LengthBucketSampler` (optional) additionally groups utterances of similar length into the same batch, which removes most of
the padding itself; it changes the batch composition the reference's random sampler would produce, so it is opt-in.
"""
import math

import torch

_ALIGN = 256   # every tensor starts on a 256-byte boundary of the flat buffer (vector loads, TMA)

# (name, dtype) of the 11 tensors in `parse_batch` order (model.py:727-753); shapes are (B,), (B, L), (B, T) or (B, M, T)
FIELDS = (('symbols', torch.int64), ('durations_float', torch.float32), ('durations_int', torch.int64),
          ('symbols_energy', torch.float32), ('symbols_pitch', torch.float32), ('input_lengths', torch.int64),
          ('frames_energy', torch.float32), ('frames_pitch', torch.float32), ('mel_specs', torch.float32),
          ('output_lengths', torch.int64), ('speaker_ids', torch.int64))


def _round_up(x, m):
    return (x + m - 1) // m * m


def _views(buf, layout):
    return tuple(buf[off:off + math.prod(shape) * torch.empty((), dtype=dt).element_size()].view(dt).view(shape)
                 for dt, shape, off in layout)


class DeviceBatch(tuple):
    """The 11 device tensors of a moved `FlatBatch`, in `DaftExprt.forward` input order: a plain tuple whose elements are views
    of ONE device byte buffer (`.flat`), so a consumer with static input buffers (graph.GraphedTrainStep) refreshes them with
    one device-to-device copy."""

    def __new__(cls, flat, layout):
        self = super().__new__(cls, _views(flat, layout))
        self.flat, self.layout = flat, layout
        return self

    def clone(self):
        return DeviceBatch(self.flat.clone(), self.layout)


class FlatBatch:
    """The 11 batch tensors as views of one byte buffer + the two file-id lists; iterates like the reference's 13-tuple."""

    def __init__(self, buf, layout, feature_dirs, feature_files):
        self.buf, self.layout = buf, layout          # layout: tuple of (dtype, shape, byte offset) per field
        self.feature_dirs, self.feature_files = feature_dirs, feature_files

    @staticmethod
    def plan(shapes):
        layout, off = [], 0
        for (_, dt), shape in zip(FIELDS, shapes):
            layout.append((dt, tuple(shape), off))
            off = _round_up(off + math.prod(shape) * torch.empty((), dtype=dt).element_size(), _ALIGN)
        return tuple(layout), off

    @classmethod
    def from_tensors(cls, tensors, feature_dirs, feature_files, shapes=None):
        """Pack (and, when `shapes` are larger than the tensors, right-zero-pad) the 11 tensors into one buffer."""
        shapes = [tuple(t.shape) for t in tensors] if shapes is None else shapes
        layout, nbytes = cls.plan(shapes)
        buf = torch.zeros(nbytes, dtype=torch.uint8)
        out = cls(buf, layout, list(feature_dirs), list(feature_files))
        for dst, src in zip(out.tensors(), tensors):
            dst[tuple(slice(0, n) for n in src.shape)].copy_(src)   # converts dtype like parse_batch's .float() / .long()
        return out

    def tensors(self):
        return _views(self.buf, self.layout)

    def key(self):
        return tuple(shape for _, shape, _ in self.layout)

    def pin_memory(self):
        """Called by DataLoader(pin_memory=True): ONE pinned allocation + copy for the whole batch."""
        return FlatBatch(self.buf.pin_memory(), self.layout, self.feature_dirs, self.feature_files)

    def to_device(self, device, non_blocking=True):
        """ONE host->device copy; returns the 11 device tensors (a `DeviceBatch`: views of the moved buffer)."""
        return DeviceBatch(self.buf.to(device, non_blocking=non_blocking), self.layout)

    # the reference unpacks the batch as a 13-tuple (model.py:729-731, fine_tune.py)
    def __iter__(self):
        return iter(self.tensors() + (self.feature_dirs, self.feature_files))

    def __len__(self):
        return 13

    def __getitem__(self, i):
        return (self.tensors() + (self.feature_dirs, self.feature_files))[i]

    @property
    def nbytes(self):
        return self.buf.numel()


class BucketedCollate:
    """collate_fn: the base collate's 13-tuple, right-padded to bucket shapes and packed into a `FlatBatch`.

    L is padded to a multiple of `l_step`, T to a multiple of `t_step` (defaults: 64 phonemes, 128 frames = one GEMM row tile);
    `l_max` / `t_max` (optional) cap the grid so that the largest bucket is exactly the dataset maximum."""

    def __init__(self, base_collate=None, l_step=64, t_step=128, l_max=None, t_max=None, flat=True):
        self.base, self.l_step, self.t_step, self.l_max, self.t_max, self.flat = base_collate, l_step, t_step, l_max, t_max, flat

    def bucket(self, L, T):
        Lb, Tb = _round_up(L, self.l_step), _round_up(T, self.t_step)
        if self.l_max is not None:
            Lb = max(L, min(Lb, self.l_max))
        if self.t_max is not None:
            Tb = max(T, min(Tb, self.t_max))
        return Lb, Tb

    def __call__(self, batch):
        if self.base is not None:
            batch = self.base(batch)
        tensors, dirs, files = tuple(batch[:11]), batch[11], batch[12]
        B, L = tensors[0].shape
        M, T = tensors[8].shape[1], tensors[8].shape[2]
        Lb, Tb = self.bucket(L, T)
        shapes = [(B, Lb)] * 5 + [(B,), (B, Tb), (B, Tb), (B, M, Tb), (B,), (B,)]
        out = FlatBatch.from_tensors(tensors, dirs, files, shapes)
        return out if self.flat else tuple(out.tensors()) + (out.feature_dirs, out.feature_files)


class LengthBucketSampler(torch.utils.data.Sampler):
    """Batch sampler: utterances are sorted by length inside shuffled chunks of `chunk_batches` batches, so each batch holds
    utterances of similar length (less padding, fewer distinct bucket shapes); batches are then shuffled.  Rank-aware like
    DistributedSampler (`rank` / `world`): every rank sees the same number of batches.  Opt-in: the reference samples uniformly
    (data_loader.py:232-239)."""

    def __init__(self, lengths, batch_size, rank=0, world=1, chunk_batches=16, seed=0, drop_last=True):
        self.lengths, self.bs, self.rank, self.world = list(lengths), batch_size, rank, world
        self.chunk, self.seed, self.drop_last, self.epoch = chunk_batches * batch_size * world, seed, drop_last, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def _batches(self):
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        order = torch.randperm(len(self.lengths), generator=g).tolist()
        batches = []
        for c0 in range(0, len(order), self.chunk):
            chunk = sorted(order[c0:c0 + self.chunk], key=lambda i: self.lengths[i])
            for b0 in range(0, len(chunk), self.bs):
                b = chunk[b0:b0 + self.bs]
                if len(b) == self.bs or not self.drop_last:
                    batches.append(b)
        n = len(batches) // self.world * self.world
        perm = torch.randperm(n, generator=g).tolist()
        return [batches[i] for i in perm][self.rank::self.world]

    def __iter__(self):
        return iter(self._batches())

    def __len__(self):
        return len(self._batches())


class BatchPrefetcher:
    def __init__(self, model, gpu):
        self.model = model.module if hasattr(model, 'module') else model
        self.gpu = gpu
        self.device = torch.device('cuda', gpu) if isinstance(gpu, int) else torch.device(gpu)
        self.stream = torch.cuda.Stream(device=self.device)
        self.pending = None

    def submit(self, batch):
        """Start the host->device copies of `batch` (the 13-tuple / FlatBatch of `parse_batch`) on the side stream."""
        assert self.pending is None, 'BatchPrefetcher: one batch in flight at a time (call get() first)'
        with torch.cuda.stream(self.stream):
            parsed = self.model.parse_batch(self.gpu, batch)
            event = torch.cuda.Event()
            event.record(self.stream)
        self.pending = (parsed, event)

    def get(self):
        """-> (inputs, targets, file_ids) of the submitted batch, ordered after its copies on the CURRENT stream."""
        assert self.pending is not None, 'BatchPrefetcher: nothing submitted'
        (inputs, targets, file_ids), event = self.pending
        self.pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(event)
        for t in tuple(inputs) + tuple(targets):
            t.record_stream(cur)   # allocated on the side stream, consumed on this one
        return inputs, targets, file_ids

    def drop(self):
        """Forget a submitted batch (its copies still complete)."""
        if self.pending is not None:
            self.pending[1].synchronize()
            self.pending = None
