"""Two-GPU (NCCL) checks of the data-parallel path — run with `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu2.py -m gpu`;
skipped on a single-GPU box.  (1) The flat gradient bucket after the NCCL all-reduce equals the mean of the ranks' local
buckets and the loss that rode along equals the mean loss; the whole step (collective included) replays as one CUDA graph
and matches the eager step.  (2) ZeRO-1: reduce-scatter / sharded clip + AdamW / all-gather with the REAL collectives
ends with the parameters of the replicated update."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gpt2_vision_language_b200 import gpt2
    from gpt2_vision_language_b200.dp import broadcast_parameters
    from gpt2_vision_language_b200.step import PretrainStep
    g = torch.load(os.path.join(GOLD, "gpt2_tiny.pt"), map_location="cpu", weights_only=False)
    out = {}

    def fresh():
        m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
        m.load_state_dict(g["sd"])
        m = m.to(dev).to(torch.bfloat16)
        broadcast_parameters(m)
        return m
    gen = torch.Generator().manual_seed(100 + rank)                 # different data per rank
    accum, mb, T = 2, 2, 24
    xs = torch.randint(0, 256, (accum, mb, T), generator=gen).to(dev)
    ys = torch.randint(0, 256, (accum, mb, T), generator=gen).to(dev)

    # ---- (1) reduced bucket = mean of local buckets; loss rides along; graph replay = eager ---------------------
    m = fresh()
    st = PretrainStep(m, micro_batch=mb, seq=T, grad_accum=accum, lr=3e-3, use_graph=False, overlap_comm=False)
    st.load_tokens(xs, ys)
    st.bucket.zero()
    st.loss.zero_()
    for i in range(accum):
        st._set_slot(i)
        st._micro()
    local = st.bucket.flat.clone()
    local_loss = st.loss.clone()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    losses = [torch.empty_like(local_loss) for _ in range(world)]
    dist.all_gather(losses, local_loss)
    st._exchange()
    mean = sum(t.float() for t in gathered) / world
    lo, hi = st.bucket.params_off, st.bucket.params_end
    err = ((st.bucket.flat[lo:hi].float() - mean[lo:hi]).abs().max() / mean[lo:hi].abs().max()).item()
    out["bucket_err"] = err
    out["differ"] = (gathered[0].float() - gathered[1].float()).abs().max().item()
    out["loss_err"] = abs(st.loss.item() - sum(t.item() for t in losses) / world)
    import gc
    for overlap in (False, True):
        res = {}
        # eager | graphs with the collectives between them (default) | ONE graph with the collectives captured inside
        for mode, (use_graph, in_graph) in {"eager": (False, False), "between": (True, False), "inside": (True, True)}.items():
            m = fresh()
            st = PretrainStep(m, micro_batch=mb, seq=T, grad_accum=accum, lr=3e-3, use_graph=use_graph, overlap_comm=overlap,
                              nccl_in_graph=in_graph)
            st.load_tokens(xs, ys)
            ls = [st.run().item() for _ in range(5)]
            torch.cuda.synchronize()
            res[mode] = (ls, torch.cat([p.detach().float().flatten() for p in m.parameters()]))
            del st, m                   # graphs that captured NCCL work must die before the process group does
            gc.collect()
        for mode in ("between", "inside"):
            out[f"{mode}_vs_eager_loss_overlap{int(overlap)}"] = max(abs(a - b) / abs(b) for a, b in zip(res[mode][0], res["eager"][0]))
            out[f"{mode}_vs_eager_w_overlap{int(overlap)}"] = (res[mode][1] - res["eager"][1]).abs().max().item()
        out[f"losses_overlap{int(overlap)}"] = res["between"][0]
        w = res["inside"][1]
        others = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(others, w)
        out[f"replicas_identical_overlap{int(overlap)}"] = bool(torch.equal(others[0], others[1]))

    # ---- (2) ZeRO-1 with real collectives vs the replicated update ------------------------------------------------
    res = {}
    for z in (False, True):
        m = fresh()
        st = PretrainStep(m, micro_batch=mb, seq=T, grad_accum=accum, lr=3e-3, use_graph=True, zero1=z, overlap_comm=False)
        st.load_tokens(xs, ys)
        ls = [st.run().item() for _ in range(4)]
        torch.cuda.synchronize()
        res[z] = (ls, st.norm.item(), torch.cat([p.detach().float().flatten() for p in m.parameters()]))
        del st, m
        gc.collect()
    out["zero1_loss"] = max(abs(a - b) / abs(b) for a, b in zip(res[True][0], res[False][0]))
    out["zero1_norm"] = abs(res[True][1] - res[False][1]) / res[False][1]
    out["zero1_w"] = (res[True][2] - res[False][2]).abs().max().item()
    q.put((rank, out))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_exchange_graph_and_zero1():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    print(res[0])
    for p in procs:
        p.join(timeout=120)
        if p.exitcode is None:
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in (0, 1):
        o = res[r]
        assert o["differ"] > 0                                   # the ranks really had different gradients
        assert o["bucket_err"] < 1e-2, o                         # bf16 ring partial sums vs fp32 mean
        assert o["loss_err"] < 1e-6, o                           # the digits are exact
        for ov in (0, 1):
            for mode in ("between", "inside"):
                assert o[f"{mode}_vs_eager_loss_overlap{ov}"] < 2e-3, o
                assert o[f"{mode}_vs_eager_w_overlap{ov}"] < 2e-2, o
            assert o[f"replicas_identical_overlap{ov}"], o
            assert o[f"losses_overlap{ov}"][-1] < o[f"losses_overlap{ov}"][0], o
        assert o["zero1_loss"] < 2e-3 and o["zero1_norm"] < 1e-2 and o["zero1_w"] < 1e-2, o
    assert res[0]["losses_overlap0"] == res[1]["losses_overlap0"]   # both ranks report the same (averaged) loss
