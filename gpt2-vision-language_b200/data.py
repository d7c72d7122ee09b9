"""Host-side input formatting either side of the path: caption ids -> (x, y, mask) exactly as
``CocoClipFullTokensDataset._encode_caption`` builds them (source/gpt2_linear/data.py:35-49), and the synthetic
batches BASELINE.json's configs are measured on.  Pure host code: no kernels, nothing from oracle/.
"""
import torch

EOT = 50256          # tiktoken gpt2 eot_token (source/gpt2_linear/train.py:96, data.py:24)


def encode_caption_ids(ids, max_len=32, eot=EOT):
    """Token ids of one caption -> x, y int64 [max_len-1], mask bool [max_len-1]  (data.py:35-49):
    empty -> [eot]; truncate to max_len-1 and append eot; pad with eot; x = ids[:-1], y = ids[1:];
    the first max(L-1, 1) targets are valid."""
    ids = list(ids)
    if len(ids) == 0:
        ids = [eot]
    ids = ids[: max_len - 1] + [eot]
    L = len(ids)
    ids = ids + [eot] * (max_len - L)
    ids = torch.tensor(ids, dtype=torch.long)
    mask = torch.zeros(max_len - 1, dtype=torch.bool)
    mask[: max(L - 1, 1)] = True
    return ids[:-1], ids[1:], mask


def collate_captions(list_of_ids, max_len=32, eot=EOT):
    """-> x, y [B,max_len-1] int64, mask [B,max_len-1] bool, labels = y.masked_fill(~mask, -100) (train.py:305-306)."""
    xs, ys, ms = zip(*(encode_caption_ids(ids, max_len, eot) for ids in list_of_ids))
    x, y, m = torch.stack(xs), torch.stack(ys), torch.stack(ms)
    return x, y, m, y.masked_fill(~m, -100)


def synthetic_caption_batch(B, seed, max_len=32, vocab=50257, eot=EOT):
    """Random captions of 8..31 ids per sample (SURVEY 8d "Synthetic inputs"), encoded like the reference does."""
    g = torch.Generator().manual_seed(seed)
    caps = []
    for _ in range(B):
        n = int(torch.randint(8, 32, (1,), generator=g))
        caps.append(torch.randint(0, vocab - 1, (n,), generator=g).tolist())
    return collate_captions(caps, max_len, eot)


def synthetic_pixels(B, seed):
    return torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(seed))


# ----------------------------------------------------------------------------------------------------------
# on-disk formats either side of the path
# ----------------------------------------------------------------------------------------------------------
class TokenShardLoader:
    """Pretraining token stream with the reference's sharding arithmetic (``DataLoaderLite``,
    source/gpt2/train_gpt2.py:149-187): ``.npy`` shards whose names contain the split, sorted; rank r starts at
    B*T*r and every batch advances all ranks by B*T*num_processes; x = buf[:-1], y = buf[1:]; the next shard is
    loaded when the following batch (+1 token) would not fit.  Batches come back as pinned int64 host tensors (or on
    ``device`` when given) so the H2D copy can overlap compute."""

    def __init__(self, B, T, process_rank, num_processes, split, data_root, device=None, pin=False):
        import os
        if split not in ("train", "val"):
            raise AssertionError("split must be 'train' or 'val'")
        self.B, self.T, self.process_rank, self.num_processes = B, T, process_rank, num_processes
        self.device, self.pin = device, pin
        shards = sorted(s for s in os.listdir(data_root) if split in s)
        self.shards = [os.path.join(data_root, s) for s in shards]
        if not self.shards:
            raise AssertionError(f"no shards found for split {split}")
        self.reset()

    @staticmethod
    def load_tokens(filename):
        import numpy as np
        return torch.from_numpy(np.load(filename).astype("int64"))     # np.int32 -> torch.long in the reference

    def reset(self):
        self.current_shard = 0
        self.tokens = self.load_tokens(self.shards[0])
        self.current_position = self.B * self.T * self.process_rank

    def next_batch(self):
        B, T = self.B, self.T
        buf = self.tokens[self.current_position: self.current_position + B * T + 1]
        x, y = buf[:-1].view(B, T), buf[1:].view(B, T)
        self.current_position += B * T * self.num_processes
        if self.current_position + (B * T * self.num_processes + 1) > len(self.tokens):
            self.current_shard = (self.current_shard + 1) % len(self.shards)
            self.tokens = self.load_tokens(self.shards[self.current_shard])
            self.current_position = B * T * self.process_rank
        if self.device is not None:
            return x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)
        if self.pin:
            return x.contiguous().pin_memory(), y.contiguous().pin_memory()
        return x, y


class ClipTokenShards:
    """The pre-computed CLIP feature cache the caption datasets read (``CocoClipFullTokensDataset``,
    source/gpt2_linear/data.py:25-28,55-63): ``index.json`` = list of ``{"shard": name, "row": r}`` (one entry per
    image, dataset order) next to ``.pt`` shards holding ``[n, 257, 768]`` tensors.  ``write`` produces that layout
    from feature batches (e.g. ``clip.ClipVisionTower`` outputs), ``__getitem__`` reads it with the reference's
    one-shard cache."""

    def __init__(self, tokens_dir):
        import json
        import os
        self.tokens_dir = tokens_dir
        with open(os.path.join(tokens_dir, "index.json")) as f:
            self.index = json.load(f)
        self._name, self._tensor = None, None

    def __len__(self):
        return len(self.index)

    def __getitem__(self, idx):
        import os
        e = self.index[idx]
        if e["shard"] != self._name:
            self._tensor = torch.load(os.path.join(self.tokens_dir, e["shard"]), map_location="cpu")
            self._name = e["shard"]
        return self._tensor[e["row"]]

    @staticmethod
    def write(tokens_dir, feature_batches, rows_per_shard=1024, dtype=torch.float16):
        """feature_batches: iterable of [b, 257, D] tensors (any device).  Returns the number of images written."""
        import json
        import os
        os.makedirs(tokens_dir, exist_ok=True)
        index, pend, n_shard = [], [], 0

        def flush():
            nonlocal pend, n_shard
            if not pend:
                return
            t = torch.cat(pend, dim=0)
            name = f"shard_{n_shard:05d}.pt"
            torch.save(t, os.path.join(tokens_dir, name))
            index.extend({"shard": name, "row": r} for r in range(t.shape[0]))
            pend, n_shard = [], n_shard + 1

        have = 0
        for fb in feature_batches:
            fb = fb.detach().to("cpu", dtype)
            while fb.shape[0]:
                take = min(rows_per_shard - have, fb.shape[0])
                pend.append(fb[:take])
                fb, have = fb[take:], have + take
                if have == rows_per_shard:
                    flush()
                    have = 0
        flush()
        with open(os.path.join(tokens_dir, "index.json"), "w") as f:
            json.dump(index, f)
        return len(index)


def save_checkpoint(path, raw_model, optimizer, step, val_loss, world_size=1):
    """The reference's rolling-checkpoint dict (source/gpt2/train_gpt2.py:363-375), written atomically."""
    import os
    import time
    tmp = path + ".tmp"
    torch.save({"model": raw_model.state_dict(), "optimizer": optimizer.state_dict() if optimizer is not None else None,
                "config": getattr(raw_model, "config", None), "step": int(step), "val_loss": float(val_loss),
                "ddp_world_size": int(world_size), "ts": time.strftime("%Y-%m-%d %H:%M:%S")}, tmp)
    os.replace(tmp, path)


def load_checkpoint(path, raw_model, optimizer=None, map_location="cpu", strict=True):
    """Resume like source/gpt2/train_gpt2.py:319-325: returns the step to start from.  A checkpoint written by the
    reference itself loads too (same state_dict keys); ``strict=False`` mirrors gpt2_linear/train.py:103-104."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    raw_model.load_state_dict(sd, strict=strict)
    if optimizer is not None and isinstance(ckpt, dict) and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    return int(ckpt.get("step", 0)) + 1 if isinstance(ckpt, dict) else 0


class CaptionFeatureDataset:
    """(x, y, mask, z) items exactly as ``CocoClipFullTokensDataset.__getitem__`` builds them
    (source/gpt2_linear/data.py:51-63): one of the image's captions is drawn at random (``random.choice``), encoded
    with ``encode_caption_ids`` and paired with the image's pre-computed CLIP tokens ``z [257, D]`` from the shard
    cache.  The COCO annotation reader and the tokenizer stay outside (control plane): ``captions[i]`` is the list of
    already-tokenised captions of image i."""

    def __init__(self, tokens_dir, captions, max_len=32, eot=EOT, seed=None):
        import random
        self.features = ClipTokenShards(tokens_dir)
        if len(self.features) != len(captions):
            raise AssertionError("index.json length mismatch with the caption list")      # data.py:29
        self.captions, self.max_len, self.eot = captions, max_len, eot
        self.rng = random.Random(seed)

    def __len__(self):
        return len(self.captions)

    def __getitem__(self, idx):
        x, y, m = encode_caption_ids(self.rng.choice(self.captions[idx]), self.max_len, self.eot)
        return x, y, m, self.features[idx]


def caption_batches(dataset, batch_size, indices=None, pin=True, drop_last=True):
    """Batches (x [B,T], y [B,T], mask [B,T] bool, z [B,257,D]) as pinned host tensors — what the train loop moves to
    the device at the top of a step (source/gpt2_linear/train.py:297-303); feed them to ``step.HostBatchFeeder`` /
    ``pool_clip_197_to_33_avg_with_cls`` or straight to a captioner."""
    idx = list(range(len(dataset))) if indices is None else list(indices)
    for s in range(0, len(idx), batch_size):
        chunk = idx[s:s + batch_size]
        if len(chunk) < batch_size and drop_last:
            return
        items = [dataset[i] for i in chunk]
        batch = tuple(torch.stack(col) for col in zip(*items))
        yield tuple(t.pin_memory() for t in batch) if pin and torch.cuda.is_available() else batch
