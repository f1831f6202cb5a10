"""Host -> device hand-over of a collated batch, one step ahead.

The reference's loop (`train.py:368-378`) calls `model.parse_batch(gpu, batch)` — eleven `.cuda(non_blocking=True)` copies —
and then the model, back to back on one stream, so the PCIe transfer of step i+1's inputs (10.7 MB at B=32, T<=1000) waits for
step i and delays step i+1.  `BatchPrefetcher` issues the SAME `parse_batch` on a side stream for the next batch while the
current step computes:

    pre = BatchPrefetcher(model, gpu)
    pre.submit(next(loader))
    for batch in loader:
        inputs, targets, file_ids = pre.get()      # waits (on the device) for the copies of THIS step
        pre.submit(batch)                          # next step's copies overlap this step's kernels
        outputs = model(inputs); ...

Batches must come from pinned memory (`DataLoader(pin_memory=True)`, as the reference does, `data_loader.py:249`) for the copies
to be asynchronous.
"""
import torch


class BatchPrefetcher:
    def __init__(self, model, gpu):
        self.model = model.module if hasattr(model, 'module') else model
        self.gpu = gpu
        self.device = torch.device('cuda', gpu) if isinstance(gpu, int) else torch.device(gpu)
        self.stream = torch.cuda.Stream(device=self.device)
        self.pending = None

    def submit(self, batch):
        """Start the host->device copies of `batch` (the 13-tuple of `parse_batch`) on the side stream."""
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
