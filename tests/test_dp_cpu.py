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
    bucket = FlatGradBucket(mod.parameters(), scalar_slot=True)
    assert len(bucket.params) == 2 and bucket.params_off == 8
    assert lin.weight.grad.data_ptr() == bucket.flat[bucket.params_off:].data_ptr()
    bucket.zero()
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(5, 24, generator=g)
    mod(x).square().mean().backward()             # autograd accumulates into the flat views
    local = bucket.flat.clone()
    # the exchange in two ranges (the split-backward overlap path reduces the tail of the bucket first) ...
    k = bucket.offset_of(lin.bias)
    assert 0 < k < bucket.params_end
    bucket.all_reduce(lo=k, hi=bucket.params_end)
    assert torch.equal(bucket.flat[:k], local[:k])            # ... leaves the head untouched until its own turn
    # the step's loss rides in the head range as exact base-16 digits (FlatGradBucket.pack_scalar)
    loss = torch.tensor(3.25 + 1.0 / 3.0 + rank * 2.7182818)
    assert bucket.scalar_rides_along()
    bucket.pack_scalar(loss)
    local[:8] = bucket.flat[:8]
    bucket.all_reduce(lo=0, hi=k)
    mean_loss = torch.zeros(())
    bucket.unpack_scalar(mean_loss)
    expect = sum(float(torch.tensor(3.25 + 1.0 / 3.0 + r * 2.7182818)) for r in range(world)) / world
    assert abs(mean_loss.item() - expect) < 2e-7, (mean_loss.item(), expect)
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
    assert torch.allclose(g0.flatten(), a0[8: 8 + g0.numel()])   # .grad is a view of the reduced bucket (after the scalar slots)


def test_scalar_digits_survive_bf16_reduction_exactly():
    """The loss travels through a bf16 AVG all-reduce as 8 base-16 digits: whatever the order of the additions (ring,
    tree, in-switch) and with every partial sum rounded to bf16, the mean is recovered to the Q8.24 resolution."""
    from gpt2_vision_language_b200.dp import FlatGradBucket
    g = torch.Generator().manual_seed(0)
    for world in (2, 4, 8, 16):
        for _ in range(20):
            losses = torch.rand(world, generator=g) * 12.0
            buckets = []
            for r in range(world):
                b = FlatGradBucket([torch.nn.Parameter(torch.zeros(4, dtype=torch.bfloat16))], scalar_slot=True)
                b.pack_scalar(losses[r])
                buckets.append(b)
            order = torch.randperm(world, generator=g).tolist()
            acc = torch.zeros(8, dtype=torch.bfloat16)
            for r in order:                                   # bf16 accumulator: rounded after every addition
                acc = (acc + buckets[r].flat[:8]).to(torch.bfloat16)
            buckets[0].flat[:8] = (acc / world).to(torch.bfloat16)
            out = torch.zeros(())
            buckets[0].unpack_scalar(out)
            expect = sum(round(float(x) * (1 << 24)) for x in losses) / world / (1 << 24)
            assert abs(out.item() - expect) < 1e-6, (world, out.item(), expect)
