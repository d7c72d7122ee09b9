"""World-size-2 gloo test of the data-parallel gradient exchange (dp.FlatGradBucket) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpt2_vision_language_b200.dp import FlatGradBucket, broadcast_parameters
    torch.manual_seed(rank)                       # different init per rank ...
    lin = torch.nn.Linear(24, 8)
    frozen = torch.nn.Linear(8, 8)
    for p in frozen.parameters():
        p.requires_grad_(False)
    mod = torch.nn.Sequential(lin, frozen)
    broadcast_parameters(mod)                     # ... made identical by the one-time broadcast
    bucket = FlatGradBucket(mod.parameters())
    assert len(bucket.params) == 2 and lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    bucket.zero()
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(5, 24, generator=g)
    mod(x).square().mean().backward()             # autograd accumulates into the flat views
    local = bucket.flat.clone()
    # the exchange in two ranges (the split-backward overlap path reduces the tail of the bucket first) ...
    k = bucket.offset_of(lin.bias)
    assert 0 < k < bucket.extra_off
    bucket.all_reduce(lo=k, hi=bucket.extra_off)
    assert torch.equal(bucket.flat[:k], local[:k])            # ... leaves the head untouched until its own turn
    bucket.all_reduce(lo=0, hi=k)
    two_step = bucket.flat.clone()
    bucket.flat.copy_(local)
    bucket.all_reduce()
    assert torch.allclose(two_step, bucket.flat, atol=1e-7)
    # numpy arrays are pickled by value: torch tensors would travel as shared-memory handles that die with the worker
    out.put((rank,) + tuple(t.detach().clone().numpy() for t in (lin.weight, local, bucket.flat, lin.weight.grad)))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0, l0, a0, g0), (_, w1, l1, a1, g1) = [(r[0],) + tuple(torch.from_numpy(a) for a in r[1:]) for r in res]
    assert torch.equal(w0, w1)                                   # broadcast made the replicas identical
    assert not torch.allclose(l0, l1)                            # different data -> different local grads
    assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-6)          # AVG, like DDP
    assert torch.equal(a0, a1)
    assert torch.allclose(g0.flatten(), a0[: g0.numel()])        # .grad is a view of the reduced bucket
