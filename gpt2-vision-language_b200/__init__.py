"""B200-native captioning / GPT-2 training step (drop-in for theophile-lt/gpt2-vision-language's modules).

Sub-modules (import lazily; none of them needs a GPU at import time):
  _lib      ctypes binding of libvlk.so (the C ABI in include/vlk.h)
  ops       tensor-level wrappers + autograd Functions over the C ABI
  gpt2      GPTConfig / CausalSelfAttention / MLP / Block / GPT / GPT_previous
  caption   Linear_Bridge, QFormerLayer, BLIP2Bridge, GPT_Caption, pool_clip_197_to_33_avg_with_cls
  xattn     CrossAttention, Vision_projector, cross-attention Block / GPT
  clip      CLIP ViT-L/14 vision tower on the same kernels
  optim     fused clip-norm + AdamW
  dp        data-parallel gradient exchange (flat bucket, NCCL side stream)
"""
__version__ = "0.1.0"
