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
