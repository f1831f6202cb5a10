"""CPU test of the multi-process plumbing: 2 ranks over gloo, flat gradient bucket, one all-reduce (mean), parameter
broadcast.  The same code path runs over NCCL on the GPU box."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from daft_exprt_b200.ddp import FlatGradSync, broadcast_parameters
    torch.manual_seed(100 + rank)                       # ranks start from different weights
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    broadcast_parameters(model, src=0)
    sync = FlatGradSync(model.parameters())
    assert sync.numel == sum(p.numel() for p in model.parameters())
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(8, 7, generator=g)              # global batch of 8 "utterances", sharded 4 + 4
    x = x_all[rank * 4:(rank + 1) * 4]
    sync.zero_grad()
    model(x).pow(2).mean().backward()
    assert sync.grads_attached()                        # autograd accumulated into the flat buffer in place
    sync.all_reduce_mean()
    # the 'gather' mode must give the same bucket without aliasing
    model2 = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    model2.load_state_dict(model.state_dict())
    sync2 = FlatGradSync(model2.parameters(), mode='gather')
    sync2.zero_grad()
    model2(x).pow(2).mean().backward()
    sync2.all_reduce_mean()
    assert torch.allclose(torch.cat([v.reshape(-1) for v in sync2.views]), torch.cat([p.grad.reshape(-1) for p in model.parameters()]))
    ret[rank] = (torch.cat([p.grad.reshape(-1) for p in model.parameters()]), torch.cat([p.data.reshape(-1) for p in model.parameters()]))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_two_ranks_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (g0, p0), (g1, p1) = ret[0], ret[1]
    assert torch.equal(p0, p1)                          # broadcast worked
    assert torch.allclose(g0, g1, atol=0, rtol=0)       # identical averaged gradient on both ranks
    # equals the single-process gradient of the mean over the two shards
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    x_all = torch.randn(8, 7, generator=torch.Generator().manual_seed(7))
    loss = 0.5 * (model(x_all[:4]).pow(2).mean() + model(x_all[4:]).pow(2).mean())
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(g0, ref, atol=1e-6)
