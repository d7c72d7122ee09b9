"""Step-level parity: a few full optimizer steps on the B200 path vs the fp32 CPU oracle — loss
trajectory and updated weights.  bf16 parameters/moments on the GPU vs fp32 on the CPU, hence the looser bars
after the first step."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def bf16_round(sd):
    return {k: v.to(torch.bfloat16).float() if v.is_floating_point() else v for k, v in sd.items()}


@pytest.mark.parametrize("use_graph", [False, True])
def test_caption_linear_steps_end_to_end(cuda, use_graph):
    """pixels -> tiny CLIP tower -> pool -> linear bridge -> GPT-2 -> CE -> backward -> clip + AdamW, 4 steps."""
    from gpt2_vision_language_b200 import gpt2, gpt2_linear
    from gpt2_vision_language_b200.clip import ClipVisionTower
    from gpt2_vision_language_b200.step import CaptionTrainStep
    from oracle import torch_oracle as O
    g, gc = load("caption_linear_tiny.pt"), load("clip_tiny.pt")
    cfg = g["cfg"]
    m = gpt2_linear.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(gpt2.GPTConfig(**cfg)), m_vis_tokens=32)
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    tower = ClipVisionTower.from_state_dict(gc["sd"], layers=2, heads=2, device=cuda)
    B, T = 2, 15
    step = CaptionTrainStep(m, tower, "linear", B, T, lr=1e-2, weight_decay=0.1, use_graph=use_graph)
    pixels = gc["pixels"].float()
    x, labels = g["input_ids"][:B], g["labels"][:B]
    mask = labels != -100
    y = labels.clamp_min(0)
    step.load_batch(pixels.to(cuda), x.to(cuda), y.to(cuda), mask.to(cuda))
    # oracle: fp32 on bf16-rounded initial weights
    sd = bf16_round(g["sd"])
    csd = bf16_round(gc["sd"])
    names = ["bridge.vis_proj.weight", "bridge.vis_proj.bias"]
    for n in names:
        sd[n].requires_grad_(True)
    mom = [torch.zeros_like(sd[n]) for n in names], [torch.zeros_like(sd[n]) for n in names]
    with torch.no_grad():
        z = O.pool33(O.clip_features(csd, pixels.to(torch.bfloat16).float(), 2, 2))
    losses, ref_losses = [], []
    for it in range(1, 5):
        losses.append(step.run().item())
        for n in names:
            sd[n].grad = None
        _, loss = O.caption_linear_forward(sd, z, x, labels, cfg["n_layer"], cfg["n_head"])
        loss.backward()
        ref_losses.append(loss.item())
        with torch.no_grad():
            O.clip_and_adamw([sd[n] for n in names], [sd[n].grad for n in names], mom[0], mom[1], it, 1e-2, [0.1, 0.0])
    assert abs(losses[0] - ref_losses[0]) / ref_losses[0] < 3e-3, (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) / b < 1.5e-2, (losses, ref_losses)
    assert losses[-1] < losses[0]                                         # it actually trains
    w = m.bridge.vis_proj.weight.detach().float().cpu()
    assert F.cosine_similarity((w - g["sd"]["bridge.vis_proj.weight"]).flatten(),
                               (sd["bridge.vis_proj.weight"].detach() - g["sd"]["bridge.vis_proj.weight"]).flatten(),
                               dim=0) > 0.98                              # same update direction after 4 steps


@pytest.mark.parametrize("use_graph", [False, True])
def test_gpt2_pretrain_steps_with_grad_accumulation(cuda, use_graph):
    from gpt2_vision_language_b200 import gpt2
    from gpt2_vision_language_b200.step import PretrainStep
    from oracle import torch_oracle as O
    g = load("gpt2_tiny.pt")
    cfg = g["cfg"]
    m = gpt2.GPT(gpt2.GPTConfig(**cfg))
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    accum, mb, T = 2, 3, 24
    step = PretrainStep(m, micro_batch=mb, seq=T, grad_accum=accum, lr=3e-3, weight_decay=0.1, use_graph=use_graph)
    gen = torch.Generator().manual_seed(5)
    xs = torch.randint(0, 256, (accum, mb, T), generator=gen)
    ys = torch.randint(0, 256, (accum, mb, T), generator=gen)
    step.load_tokens(xs.to(cuda), ys.to(cuda))
    sd = bf16_round(g["sd"])
    names = [n for n, _ in m.named_parameters()]                     # tied wte/lm_head appears once
    params = {n: sd[n].clone().requires_grad_(True) for n in names}
    wds = [0.1 if params[n].dim() >= 2 else 0.0 for n in names]
    mom = [torch.zeros_like(params[n]) for n in names], [torch.zeros_like(params[n]) for n in names]

    def full_sd():
        d = dict(sd)
        d.update(params)
        d["lm_head.weight"] = params["transformer.wte.weight"] if "transformer.wte.weight" in params else params["lm_head.weight"]
        d["transformer.wte.weight"] = d["lm_head.weight"]
        return d
    losses, ref_losses = [], []
    for it in range(1, 4):
        losses.append(step.run().item())
        for p in params.values():
            p.grad = None
        tot = 0.0
        for i in range(accum):
            _, loss = O.gpt2_forward(full_sd(), xs[i], ys[i], cfg["n_layer"], cfg["n_head"])
            (loss / accum).backward()
            tot += loss.item() / accum
        ref_losses.append(tot)
        with torch.no_grad():
            O.clip_and_adamw([params[n] for n in names], [params[n].grad for n in names], mom[0], mom[1], it, 3e-3, wds)
    assert abs(losses[0] - ref_losses[0]) / ref_losses[0] < 3e-3, (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) / b < 2e-2, (losses, ref_losses)
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("use_graph", [False, True])
def test_split_backward_overlap_path_matches_monolithic(cuda, use_graph):
    """The data-parallel overlap path (backward cut in the middle of the stack, upper gradients exchanged on a second
    stream under the lower half) forced on a single GPU — the collectives are no-ops, the gradient math is what is
    checked: same losses, gradient norms and updated weights as the monolithic step."""
    from gpt2_vision_language_b200 import gpt2, gpt2_cross_att as xa
    from gpt2_vision_language_b200.clip import ClipVisionTower
    from gpt2_vision_language_b200.step import CaptionTrainStep, PretrainStep
    # ---- GPT-2 pretraining step with gradient accumulation
    g = load("gpt2_tiny.pt")
    gen = torch.Generator().manual_seed(11)
    accum, mb, T = 3, 2, 24
    xs = torch.randint(0, 256, (accum, mb, T), generator=gen).to(cuda)
    ys = torch.randint(0, 256, (accum, mb, T), generator=gen).to(cuda)
    out = {}
    for overlap in (False, True):
        m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
        m.load_state_dict(g["sd"])
        m = m.to(cuda).to(torch.bfloat16)
        st = PretrainStep(m, micro_batch=mb, seq=T, grad_accum=accum, lr=3e-3, use_graph=use_graph, overlap_comm=overlap)
        assert st.overlap == overlap
        st.load_tokens(xs, ys)
        losses = [st.run().item() for _ in range(4)]
        out[overlap] = (losses, st.norm.item(), torch.cat([p.detach().float().flatten() for p in m.parameters()]))
    assert out[True][0] == pytest.approx(out[False][0], rel=2e-3)
    assert out[True][1] == pytest.approx(out[False][1], rel=2e-2)
    assert F.cosine_similarity(out[True][2], out[False][2], dim=0) > 0.9999
    assert (out[True][2] - out[False][2]).abs().max().item() < 2e-2
    if out[True][2].numel():
        lo, hi = st.upper
        assert 0 < lo < hi == st.bucket.params_end     # the upper half is a proper, non-empty tail of the bucket
    # ---- cross-attention captioning step
    g, gc = load("caption_xattn_tiny.pt"), load("clip_tiny.pt")
    B = 2
    out = {}
    for overlap in (False, True):
        m = xa.GPT(xa.GPTConfig(**g["cfg"]))
        m.load_state_dict(g["sd"])
        with torch.no_grad():
            for i, blk in enumerate(m.transformer.h):
                blk.cross_gate.fill_(0.3 + 0.1 * i)
        m = m.to(cuda).to(torch.bfloat16)
        tower = ClipVisionTower.from_state_dict(gc["sd"], layers=2, heads=2, device=cuda)
        T = g["idx"].shape[1]
        st = CaptionTrainStep(m, tower, "xattn", B, T, lr=1e-2, use_graph=use_graph, overlap_comm=overlap)
        assert st.overlap == overlap
        st.load_batch(gc["pixels"].float()[:B].to(cuda), g["idx"][:B].to(cuda), g["targets"][:B].to(cuda),
                      g["mask"][:B].to(cuda))
        losses = [st.run().item() for _ in range(4)]
        out[overlap] = (losses, st.norm.item(),
                        torch.cat([p.detach().float().flatten() for p in m.parameters() if p.requires_grad]))
    assert out[True][0] == pytest.approx(out[False][0], rel=2e-3)
    assert out[True][1] == pytest.approx(out[False][1], rel=2e-2)
    assert (out[True][2] - out[False][2]).abs().max().item() < 2e-2
    assert out[False][0][-1] < out[False][0][0]


def test_zero1_update_matches_full_update(cuda):
    """ZeRO-1 (reduce-scatter / sharded clip+AdamW / all-gather) against the replicated FusedAdamW on one GPU:
    (a) PretrainStep(zero1=True) at world size 1 reproduces the regular step; (b) three FAKE ranks, each owning a
    third of the flat bucket, driven phase by phase (the collectives replaced by local copies), end with exactly the
    parameters of the replicated update — shard boundaries cut through tensors with different weight decay."""
    from gpt2_vision_language_b200 import gpt2
    from gpt2_vision_language_b200.dp import FlatGradBucket, FlatParamBucket
    from gpt2_vision_language_b200.optim import Zero1AdamW
    from gpt2_vision_language_b200.step import PretrainStep
    g = load("gpt2_tiny.pt")
    gen = torch.Generator().manual_seed(3)
    accum, mb, T = 2, 2, 24
    xs = torch.randint(0, 256, (accum, mb, T), generator=gen).to(cuda)
    ys = torch.randint(0, 256, (accum, mb, T), generator=gen).to(cuda)

    def fresh():
        m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
        m.load_state_dict(g["sd"])
        return m.to(cuda).to(torch.bfloat16)
    # (a)
    out = {}
    for z in (False, True):
        m = fresh()
        st = PretrainStep(m, micro_batch=mb, seq=T, grad_accum=accum, lr=3e-3, use_graph=True, zero1=z)
        st.load_tokens(xs, ys)
        losses = [st.run().item() for _ in range(4)]
        out[z] = (losses, st.norm.item(), torch.cat([p.detach().float().flatten() for p in m.parameters()]))
    assert out[True][0] == pytest.approx(out[False][0], rel=2e-3)
    assert out[True][1] == pytest.approx(out[False][1], rel=1e-2)
    assert (out[True][2] - out[False][2]).abs().max().item() < 1e-2
    # (b) identical gradients on three fake ranks
    world = 3
    ref = fresh()
    ref_bucket = FlatGradBucket(ref.parameters())
    ref_opt = ref.configure_optimizers(0.1, 3e-3, "cuda")
    m = fresh()
    bucket = FlatGradBucket(m.parameters(), pad_multiple=8 * world)
    pb = FlatParamBucket(bucket)
    wds = [0.1 if p.dim() >= 2 else 0.0 for p in bucket.params]
    ranks = [Zero1AdamW(bucket, pb, wds, lr=3e-3, rank=r, world=world) for r in range(world)]
    assert len({(z.lo, z.hi) for z in ranks}) == world and sum(len(z.segments) for z in ranks) >= len(bucket.params)
    for it in range(3):
        ref_bucket.zero()
        _, loss = ref(xs[it % accum], ys[it % accum])
        loss.backward()
        # the same gradients for both updates (LayerNorm / embedding gradients use fp32 atomics: two backward passes
        # agree only to the last bit, which is not what is under test here)
        bucket.zero()
        bucket.flat[: ref_bucket.flat.numel()].copy_(ref_bucket.flat)
        norm_ref = ref_opt.clip_grad_norm(1.0)
        ref_opt.step()
        total = torch.zeros(1, device=cuda)
        for z in ranks:                       # reduce-scatter of identical gradients = take the slice
            z.gshard.copy_(bucket.flat[z.lo:z.hi])
            total += z.local_sumsq()
        assert total.sqrt().item() == pytest.approx(norm_ref.item(), rel=1e-4)
        for z in ranks:
            z.apply(total, 1.0)                # the all-gather is implicit: all fake ranks share the flat buffer
        with torch.no_grad():                  # keep the two replicas' weights identical for the next forward
            worst, differ, count = 0.0, 0, 0
            for (n1, p1), (n2, p2) in zip(ref.named_parameters(), m.named_parameters()):
                assert n1 == n2
                d = (p1.float() - p2.float()).abs()
                worst, differ, count = max(worst, d.max().item()), differ + int((d > 0).sum()), count + d.numel()
            # same arithmetic per element; only the summation order of the global norm differs (clip factor +-1e-7),
            # which can move a value across a bf16 rounding boundary once in a while
            assert worst < 5e-4 and differ / count < 1e-3, (worst, differ, count)


def test_optimizer_resume_from_cpu_mapped_checkpoint(cuda, tmp_path):
    """save_checkpoint -> load_checkpoint(map_location='cpu') -> load_state_dict -> step(): the step counter the AdamW
    kernel dereferences must live on the parameters' device again, and the resumed run must continue exactly like the
    uninterrupted one (reference resume: source/gpt2/train_gpt2.py:319-325)."""
    from gpt2_vision_language_b200 import data, gpt2
    g = load("gpt2_tiny.pt")

    def fresh():
        m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
        m.load_state_dict(g["sd"])
        return m.to(cuda).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(9)
    xs = torch.randint(0, 256, (4, 3, 24), generator=gen).to(cuda)
    ys = torch.randint(0, 256, (4, 3, 24), generator=gen).to(cuda)

    def one_step(m, opt, i):
        opt.zero_grad(set_to_none=True)
        _, loss = m(xs[i], ys[i])
        loss.backward()
        opt.clip_grad_norm(1.0)
        opt.step()
        return loss.item()
    a = fresh()
    opt_a = a.configure_optimizers(0.1, 3e-3, "cuda")
    for i in range(2):
        one_step(a, opt_a, i)
    path = str(tmp_path / "ckpt.pt")
    data.save_checkpoint(path, a, opt_a, step=2, val_loss=1.0)
    b = fresh()
    opt_b = b.configure_optimizers(0.1, 3e-3, "cuda")
    start = data.load_checkpoint(path, b, opt_b)            # default map_location = 'cpu'
    assert start == 3
    assert opt_b._step_t.is_cuda and opt_b._step_t.item() == 2.0
    la = [one_step(a, opt_a, i) for i in (2, 3)]
    lb = [one_step(b, opt_b, i) for i in (2, 3)]
    assert la == pytest.approx(lb, rel=1e-3)
    wa = torch.cat([p.detach().float().flatten() for p in a.parameters()])
    wb = torch.cat([p.detach().float().flatten() for p in b.parameters()])
    assert (wa - wb).abs().max().item() < 2e-3


def test_learning_rate_and_weight_decay_changes_reach_the_captured_update(cuda):
    """param_groups[i]['lr'] = x (train_gpt2.py:474-475) must change what a REPLAYED CUDA graph does, and a changed
    weight_decay must rebuild the device table it is baked into."""
    from gpt2_vision_language_b200 import gpt2
    from gpt2_vision_language_b200.step import PretrainStep
    g = load("gpt2_tiny.pt")
    m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    st = PretrainStep(m, micro_batch=2, seq=24, grad_accum=1, lr=1e-3, use_graph=True)
    gen = torch.Generator().manual_seed(2)
    st.load_tokens(torch.randint(0, 256, (1, 2, 24), generator=gen).to(cuda), torch.randint(0, 256, (1, 2, 24), generator=gen).to(cuda))
    w = m.transformer.h[0].mlp.c_fc.weight

    def delta():
        before = w.detach().float().clone()
        st.run()
        torch.cuda.synchronize()
        return (w.detach().float() - before).abs().mean().item()
    for _ in range(3):                       # eager warm-up, capture, first replay
        delta()
    d1 = delta()
    st.set_lr(0.0)
    before = w.detach().clone()
    st.run()
    torch.cuda.synchronize()
    assert torch.equal(before, w.detach())                     # lr = 0 under graph replay: the weights do not move
    st.set_lr(4e-3)
    assert delta() > 2.0 * d1                                  # 4 x lr: Adam's normalised update scales with lr
    # weight decay is part of the optimizer's device table
    opt = st.opt
    keys0 = opt._table_key
    opt.param_groups[0]["weight_decay"] = 0.5
    opt.clip_grad_norm(1.0)
    assert opt._table_key != keys0


def test_out_of_range_ids_and_labels_poison_the_loss(cuda):
    """torch raises device asserts for an out-of-vocabulary token id / label; the B200 path makes the loss NaN instead of
    reading out of bounds.  Only -100 is the ignore_index."""
    from gpt2_vision_language_b200 import gpt2, ops
    g = load("gpt2_tiny.pt")
    m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    idx, tgt = g["idx"].to(cuda), g["targets"].to(cuda)
    with torch.no_grad():
        ok = m(idx, tgt)[1].item()
        bad_label = tgt.clone()
        bad_label[0, 0] = 256                                   # == vocab_size
        assert torch.isnan(m(idx, bad_label)[1]).item()
        neg_label = tgt.clone()
        neg_label[0, 0] = -5                                    # negative but not ignore_index
        assert torch.isnan(m(idx, neg_label)[1]).item()
        ign = tgt.clone()
        ign[0, :5] = -100
        v = m(idx, ign)[1].item()
        assert v == v and abs(v - ok) < 0.5
        bad_id = idx.clone()
        bad_id[1, 3] = 999
        assert torch.isnan(m(bad_id, tgt)[1]).item()
    rows = ops.cross_entropy_rows(torch.randn(4, 256, device=cuda).bfloat16(), torch.tensor([1, -100, 256, 7], device=cuda))
    assert rows[1].item() == 0.0 and torch.isnan(rows[2]).item() and rows[0].item() > 0
