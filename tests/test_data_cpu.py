"""Host-side data formats either side of the path: token shards with DataLoaderLite's
arithmetic, the CLIP feature cache layout, the checkpoint dict.  No GPU, no kernels."""
import json
import os

import numpy as np
import torch

from gpt2_vision_language_b200 import data


def _reference_stream(tokens, B, T, rank, world, n_batches):
    """Restatement of DataLoaderLite.next_batch for ONE shard (train_gpt2.py:176-187)."""
    pos, out = B * T * rank, []
    for _ in range(n_batches):
        buf = tokens[pos:pos + B * T + 1]
        out.append((buf[:-1].view(B, T), buf[1:].view(B, T)))
        pos += B * T * world
        if pos + (B * T * world + 1) > len(tokens):
            pos = B * T * rank
    return out


def test_token_shard_loader_matches_reference_arithmetic(tmp_path):
    rng = np.random.default_rng(0)
    for i, split in enumerate(["train", "train", "val"]):
        np.save(tmp_path / f"edufineweb_{split}_{i:06d}.npy", rng.integers(0, 50257, size=1000 + 37 * i, dtype=np.uint16))
    B, T, world = 2, 8, 2
    for rank in range(world):
        ld = data.TokenShardLoader(B, T, rank, world, "train", str(tmp_path))
        assert len(ld.shards) == 2 and all("train" in s for s in ld.shards)
        first = torch.from_numpy(np.load(ld.shards[0]).astype("int64"))
        ref = _reference_stream(first, B, T, rank, world, 5)
        for xr, yr in ref:
            x, y = ld.next_batch()
            assert x.dtype == torch.int64 and x.shape == (B, T)
            assert torch.equal(x, xr) and torch.equal(y, yr)
            assert torch.equal(x.flatten()[1:], y.flatten()[:-1])
    # shard roll-over: 1000 tokens, B*T*world = 32: after n batches the next one needs 32n + 33 <= 1000, so the
    # loader moves to the second shard after the 31st batch
    ld = data.TokenShardLoader(B, T, 0, world, "train", str(tmp_path))
    for _ in range(30):
        ld.next_batch()
    assert ld.current_shard == 0
    ld.next_batch()
    assert ld.current_shard == 1 and ld.current_position == 0
    ranks = [data.TokenShardLoader(B, T, r, world, "train", str(tmp_path)).next_batch()[0] for r in range(world)]
    assert not torch.equal(ranks[0], ranks[1])      # ranks read disjoint slices


def test_token_shard_loader_matches_reference_dataloaderlite_golden(tmp_path):
    """Against batches drawn from the reference's OWN DataLoaderLite (train_gpt2.py:148-187 exec'd by
    oracle/make_golden.py over the same shard files): 70 batches per rank and split, so both train shards roll over
    and the stream wraps around."""
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataloader_lite.pt"),
                   weights_only=False)
    for name, toks in zip(g["names"], g["shards"]):
        np.save(tmp_path / name, toks.numpy().astype(np.uint16))
    for (split, rank), ref in g["batches"].items():
        ld = data.TokenShardLoader(g["B"], g["T"], rank, g["world"], split, str(tmp_path))
        for i in range(ref["x"].shape[0]):
            x, y = ld.next_batch()
            assert torch.equal(x, ref["x"][i].long()) and torch.equal(y, ref["y"][i].long()), (split, rank, i)
            assert ld.current_shard == int(ref["shard_after"][i]), (split, rank, i)


def test_clip_token_shards_roundtrip(tmp_path):
    g = torch.Generator().manual_seed(1)
    batches = [torch.randn(b, 257, 16, generator=g) for b in (5, 3, 7)]
    n = data.ClipTokenShards.write(str(tmp_path), batches, rows_per_shard=4, dtype=torch.float32)
    assert n == 15
    index = json.load(open(tmp_path / "index.json"))
    assert index[0] == {"shard": "shard_00000.pt", "row": 0} and index[5] == {"shard": "shard_00001.pt", "row": 1}
    assert sorted(os.listdir(tmp_path)) == ["index.json"] + [f"shard_{i:05d}.pt" for i in range(4)]
    ds = data.ClipTokenShards(str(tmp_path))
    allrows = torch.cat(batches)
    assert len(ds) == 15
    for i in (0, 3, 4, 14, 7):
        assert torch.equal(ds[i], allrows[i]) and ds[i].shape == (257, 16)


def test_checkpoint_dict_layout_and_resume(tmp_path):
    m = torch.nn.Linear(4, 3)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    m(torch.randn(2, 4)).sum().backward()
    opt.step()
    path = str(tmp_path / "model_last.pt")
    data.save_checkpoint(path, m, opt, step=41, val_loss=3.25, world_size=8)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"model", "optimizer", "config", "step", "val_loss", "ddp_world_size", "ts"}   # train_gpt2.py:365-373
    m2 = torch.nn.Linear(4, 3)
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-3)
    assert data.load_checkpoint(path, m2, opt2) == 42
    assert torch.equal(m2.weight, m.weight)
    assert opt2.state_dict()["state"][0]["step"] == opt.state_dict()["state"][0]["step"]


def test_samplers_follow_the_reference_rules():
    """Nucleus / top-k samplers of the decode path (data.py:114-125, train_gpt2.py:444-449) on a hand-made
    distribution: tokens outside the nucleus / the top-k are never drawn, the argmax always can be."""
    from gpt2_vision_language_b200.decode import _sample_top_k, _sample_top_p
    g = torch.Generator().manual_seed(0)
    logits = torch.log(torch.tensor([[0.5, 0.3, 0.15, 0.04, 0.01]])).repeat(2000, 1)
    draws = _sample_top_p(logits, 1.0, 0.9, g)
    # cumulative mass 0.5, 0.8, 0.95: the third token is the first to push the mass over 0.9 and is kept
    assert set(draws.tolist()) == {0, 1, 2}
    assert (draws == 0).float().mean().item() == __import__("pytest").approx(0.5 / 0.95, abs=0.04)
    draws = _sample_top_k(logits, 2, g)
    assert set(draws.tolist()) == {0, 1}
    one = _sample_top_p(torch.tensor([[10.0, 0.0, 0.0]]), 0.8, 0.9, g)
    assert one.tolist() == [0]


def test_auto_split_k_heuristic(monkeypatch):
    """Slice count of the deterministic split-K products (ops.auto_split_k) on a 148-SM part: weight gradients of the
    pretraining step are cut so that one wave of 74 tile slots is (nearly) full, big outputs are left alone."""
    from gpt2_vision_language_b200 import ops
    monkeypatch.setattr(ops, "_SM_COUNT", 148)
    assert ops.auto_split_k(768, 768, 16384) == 8        # 9 tiles -> 72 units
    assert ops.auto_split_k(3072, 768, 16384) == 2       # 36 tiles -> 72 units
    assert ops.auto_split_k(512, 768, 50304) == 12       # d h of a 512-row lm_head chunk: 6 tiles -> 72 units
    assert ops.auto_split_k(16384, 768, 50304) == 1      # 192 tiles already cover the GPU
    assert ops.auto_split_k(50304, 768, 16384) == 1
    assert ops.auto_split_k(768, 768, 1024) == 1         # short contraction: not worth a second pass


def test_zero1_shard_segments_cover_every_element_once():
    """dp.shard_segments: the N slices of the flat bucket, intersected with the tensor layout, tile every tensor
    exactly once and carry that tensor's weight decay."""
    from gpt2_vision_language_b200.dp import shard_segments
    numels = [50304 * 8, 1024 * 8, 7, 24, 2304 * 8, 1]
    wds = [0.1, 0.1, 0.0, 0.0, 0.1, 0.0]
    offs, total = [], 0
    for n in numels:
        offs.append(total)
        total += (n + 7) // 8 * 8
    for world in (1, 2, 3, 8):
        padded = (total + 8 * world - 1) // (8 * world) * (8 * world)
        S = padded // world
        cover = torch.zeros(padded, dtype=torch.int32)
        wd_of = torch.full((padded,), -1.0)
        for r in range(world):
            for a, n, wd in shard_segments(offs, numels, wds, r * S, (r + 1) * S):
                assert r * S <= a and a + n <= (r + 1) * S
                cover[a:a + n] += 1
                wd_of[a:a + n] = wd
        for o, n, wd in zip(offs, numels, wds):
            assert (cover[o:o + n] == 1).all() and (wd_of[o:o + n] == wd).all()
        assert int(cover.sum()) == sum(numels)            # padding belongs to no segment


def test_caption_feature_dataset_items_and_batches(tmp_path):
    """CocoClipFullTokensDataset.__getitem__ semantics on the shard cache: a random caption of the image, encoded with
    the reference rule, next to that image's [257, D] tokens; batches stack them."""
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(6, 257, 8, generator=g)
    data.ClipTokenShards.write(str(tmp_path), [feats], rows_per_shard=4, dtype=torch.float32)
    captions = [[[10 * i + k, 7, 8] + [9] * k for k in range(5)] for i in range(6)]
    ds = data.CaptionFeatureDataset(str(tmp_path), captions, max_len=8, seed=0)
    assert len(ds) == 6
    seen = set()
    for _ in range(30):
        x, y, m, z = ds[3]
        assert torch.equal(z, feats[3]) and x.shape == y.shape == m.shape == (7,)
        first = int(x[0])
        assert first in {30 + k for k in range(5)}                    # one of image 3's five captions
        seen.add(first)
        L = 3 + (first - 30) + 1                                      # caption length + EOT
        assert m.sum().item() == max(min(L, 8) - 1, 1) and torch.equal(x[1:], y[:-1])
    assert len(seen) > 1                                              # random.choice over the captions
    batches = list(data.caption_batches(ds, 4, pin=False))
    assert len(batches) == 1                                          # drop_last
    x, y, m, z = batches[0]
    assert x.shape == (4, 7) and m.dtype == torch.bool and z.shape == (4, 257, 8) and torch.equal(z, feats[:4])
    try:
        data.CaptionFeatureDataset(str(tmp_path), captions[:5])
        raise SystemExit("length mismatch not detected")
    except AssertionError:
        pass
