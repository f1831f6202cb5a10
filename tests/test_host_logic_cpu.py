"""CPU tests of the host-side callers of the path: bucketed collate / FlatBatch (row N2), sharded validation over 2 gloo ranks
(row N4), the reference learning-rate schedule, the numpy restatement of the dropout hash against known answers, and the
oracle's dropout hook."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import daft_exprt_oracle as oracle
from daft_exprt_b200 import synthetic
from daft_exprt_b200.data import BucketedCollate, FlatBatch, LengthBucketSampler
from helpers import DropoutReplay, attn_dropout_scale, hash_u32, rows_dropout_scale


def _batch13(B, L, T, seed=0):
    return tuple(synthetic.make_batch(B, L, T, 11, seed=seed)) + (['dir'] * B, [f'f{i}' for i in range(B)])


def test_bucketed_collate_pads_to_grid_and_keeps_values():
    b = _batch13(4, 37, 211)
    col = BucketedCollate(None, l_step=64, t_step=128)
    fb = col(b)
    assert isinstance(fb, FlatBatch) and len(fb) == 13
    t = fb.tensors()
    assert t[0].shape == (4, 64) and t[8].shape == (4, 80, 256) and t[6].shape == (4, 256) and t[5].shape == (4,)
    for got, ref in zip(t, b[:11]):
        assert got.dtype == (torch.int64 if not ref.is_floating_point() else torch.float32)
        sl = tuple(slice(0, n) for n in ref.shape)
        assert torch.equal(got[sl], ref.to(got.dtype))
        pad = got.clone()
        pad[sl] = 0
        assert float(pad.abs().sum()) == 0.0                      # padding is zeros, like the reference collate (data_loader.py:160-165)
    # unpacks like the reference's 13-tuple
    (sym, _, _, _, _, in_len, _, _, mel, out_len, spk, dirs, files) = fb
    assert torch.equal(in_len, b[5]) and torch.equal(out_len, b[9]) and dirs == b[11] and files == b[12]
    # one buffer, 256-byte aligned fields; the same padded shapes give the same layout key (= graph key)
    assert all(off % 256 == 0 for _, _, off in fb.layout)
    assert col(_batch13(4, 60, 250, seed=3)).key() == fb.key()
    assert col(_batch13(4, 65, 250, seed=3)).key() != fb.key()
    # caps: the largest bucket is the dataset maximum
    capped = BucketedCollate(None, 64, 128, l_max=200, t_max=1000)
    assert capped.bucket(199, 999) == (200, 1000) and capped.bucket(10, 100) == (64, 128)


def test_length_bucket_sampler_covers_dataset_once_per_rank_split():
    lengths = [int(x) for x in torch.randint(100, 1000, (203,), generator=torch.Generator().manual_seed(1))]
    seen = []
    for r in range(2):
        s = LengthBucketSampler(lengths, 8, rank=r, world=2, chunk_batches=4, seed=5)
        batches = list(s)
        assert len(batches) == len(s) and all(len(b) == 8 for b in batches)
        seen += [i for b in batches for i in b]
        spread = np.mean([max(lengths[i] for i in b) - min(lengths[i] for i in b) for b in batches])
        assert spread < 250                                      # similar lengths share a batch (uniform sampling: ~800)
    assert len(seen) == len(set(seen))                           # no utterance twice across ranks


def test_reference_lr_schedule_matches_train_py():
    from daft_exprt_b200.graph import reference_lr_schedule
    from daft_exprt_b200.hparams import default_hparams
    hp = default_hparams()
    f = reference_lr_schedule(hp)
    assert f(0) == hp.initial_learning_rate
    assert abs(f(hp.warmup_steps) - hp.max_learning_rate) < 1e-12
    assert abs(f(4 * hp.warmup_steps) - hp.max_learning_rate / 2) < 1e-12      # inverse square root decay (train.py:148)
    assert abs(f(5000) - (hp.initial_learning_rate + hp.max_learning_rate) / 2) < 1e-12


# -- sharded validation over gloo with a stub model (the arithmetic of the shard / all-reduce, not the kernels) -----------------
class _StubModel(torch.nn.Module):
    def parse_batch(self, gpu, batch):
        return batch, batch, None

    def forward(self, inputs):
        return inputs


class _StubCriterion:
    def forward_device(self, outputs, targets, iteration):
        v = outputs.float()
        return torch.stack([v * k for k in range(7)] + [v * 10.0])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _val_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from daft_exprt_b200.training import validate_sharded
    batches = [torch.tensor(float(i + 1)) for i in range(5)]
    loss, indiv, tg, out = validate_sharded('cpu', _StubModel(), _StubCriterion(), batches)
    ret[rank] = (loss, indiv, len(out))
    dist.destroy_process_group()


def test_validate_sharded_two_ranks_gloo():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_val_worker, args=(world, port, ret), nprocs=world, join=True)
    mean = sum(range(1, 6)) / 5.0
    for r in range(world):
        loss, indiv, n_local = ret[r]
        assert abs(loss - 10.0 * mean) < 1e-9                    # mean over ALL 5 batches on every rank
        assert abs(indiv['duration_loss'] - 2.0 * mean) < 1e-9 and abs(indiv['mel_spec_l2_loss'] - 6.0 * mean) < 1e-9
        assert set(indiv) == {'duration_loss', 'energy_loss', 'pitch_loss', 'mel_spec_l1_loss', 'mel_spec_l2_loss'}
    assert ret[0][2] == 3 and ret[1][2] == 2                     # rank 0 ran batches 0, 2, 4; rank 1 ran 1, 3


# -- dropout hash restatement -----------------------------------------------------------------------------------------------
def test_dropout_hash_known_answers_and_statistics():
    # known answers computed by hand from common.cuh's definition (python ints, independent of the numpy vectorisation)
    def ref_hash(seed, idx):
        m = 0xFFFFFFFF
        x = ((idx & m) * 0x9E3779B1 + (seed & m)) & m
        x ^= ((idx >> 32) * 0x85EBCA77 + (seed >> 32)) & m
        x ^= x >> 16; x = (x * 0x85EBCA6B) & m
        x ^= x >> 13; x = (x * 0xC2B2AE35) & m
        x ^= x >> 16
        return x
    seed = 0x123456789ABCDEF1
    idx = np.array([0, 1, 77, 2 ** 32 + 5, 2 ** 40 + 123], dtype=np.uint64)
    assert [int(v) for v in hash_u32(seed, idx)] == [ref_hash(seed, int(i)) for i in idx]
    s = rows_dropout_scale(999, 200000, 0.1)
    assert set(np.unique(s)) == {np.float32(0.0), np.float32(1.0) / (np.float32(1.0) - np.float32(0.1))}
    assert abs((s > 0).mean() - 0.9) < 5 * np.sqrt(0.09 / s.size)
    a = attn_dropout_scale(4242, 2, 3, 257, 0.25)
    assert a.shape == (2, 3, 257, 257) and abs((a > 0).mean() - 0.75) < 5 * np.sqrt(0.1875 / a.size)


def test_oracle_dropout_hook_call_order_and_identity():
    """A hook that keeps everything reproduces the eval forward; the number of sites matches the product's seed draws:
    3 (pre-net) + 12 blocks x 3 + 2 (predictor) = 41."""
    class KeepAll:
        n = 0

        def rows(self, x, p):
            KeepAll.n += 1
            return x

        def attn(self, probs, p):
            KeepAll.n += 1
            return probs
    from daft_exprt_b200.model import reference_state_shapes
    sd = synthetic.synthetic_state_dict(reference_state_shapes(12), 1234)
    ohp = oracle.OracleHParams(n_speakers=12)
    inputs = synthetic.make_batch(2, 9, 31, 11, seed=5)
    with torch.no_grad():
        ref = oracle.forward(sd, ohp, inputs)
        got = oracle.forward(sd, ohp, inputs, dropout=KeepAll())
    assert KeepAll.n == 41
    assert torch.equal(ref[3][0], got[3][0])
    seeds = list(range(1, 42))
    with torch.no_grad():
        dropped = oracle.forward(sd, ohp, inputs, dropout=DropoutReplay(seeds))
    assert not torch.equal(dropped[3][0], ref[3][0]) and torch.isfinite(dropped[3][0]).all()
