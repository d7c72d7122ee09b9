"""tcgen05 GEMM (vlk_gemm_bf16) against a torch fp32 matmul on the same bf16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_act(x, act):
    import torch.nn.functional as F
    if act == "gelu_tanh":
        return F.gelu(x, approximate="tanh")
    if act == "gelu_erf":
        return F.gelu(x)
    if act == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    return x


def _check(out, ref, tol=2e-2):
    out = out.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err / scale < tol, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 128), (4096, 768, 768), (16448, 1024, 1024),
                                   (1984, 2304, 768), (2112, 768, 768), (300, 64, 192), (77, 264, 72),
                                   (512, 50304, 768)])
@pytest.mark.parametrize("bn", [None, 256, 128, 64])
def test_gemm_nt_plain(cuda, M, N, K, bn):
    from gpt2_vision_language_b200 import ops
    if bn is not None and M * N > 20e6:
        pytest.skip("forced-tile sweep only on small/medium shapes")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = torch.randn(N, K, device=cuda, generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    if bn is None:
        _check(ops.gemm(a, b), ref)
    else:                      # every tile width, as a CTA pair (cta_group::2) and as single CTAs
        for pair in ((0, 1) if bn >= 128 else (0,)):
            _check(ops.gemm_tile(a, b, tile_n=bn, cta_pair=pair), ref)


@pytest.mark.parametrize("act", ["gelu_tanh", "gelu_erf", "quick_gelu", None])
def test_gemm_epilogues(cuda, act):
    from gpt2_vision_language_b200 import ops
    M, N, K = 1000, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g).bfloat16()
    res = torch.randn(M, N, device=cuda, generator=g).bfloat16()
    scale = torch.tensor([0.37], device=cuda)
    out, aux = ops.gemm(a, w, bias=bias, residual=res, act=act, aux_out=True, scale=scale, alpha=0.5)
    pre = 0.5 * (a.float() @ w.float().t()) + bias.float()
    _check(aux, pre)
    ref = _ref_act(pre, act) * 0.37 + res.float()
    _check(out, ref)
    # activation-gradient epilogue: D = acc * act'(aux)
    if act is not None:
        dy = torch.randn(M, K, device=cuda, generator=g).bfloat16()
        wt = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
        d = ops.gemm(dy, wt, aux_in=aux, act=act, dact=True)
        x = aux.float().requires_grad_(True)
        _ref_act(x, act).backward(dy.float() @ wt.float().t())
        _check(d, x.grad)


@pytest.mark.parametrize("trans_a,trans_b", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (768, 768, 2112), (4096, 768, 2304), (1984, 768, 50304)])
def test_gemm_transposed_operands(cuda, trans_a, trans_b, M, N, K):
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn((K, M) if trans_a else (M, K), device=cuda, generator=g).bfloat16()
    b = torch.randn((K, N) if trans_b else (N, K), device=cuda, generator=g).bfloat16()
    out = ops.gemm(a, b, trans_a=trans_a, trans_b=trans_b)
    A = a.float().t() if trans_a else a.float()
    Bm = b.float() if trans_b else b.float().t()
    _check(out, A @ Bm)


def test_gemm_fp32_out_and_strided(cuda):
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn(640, 2304, device=cuda, generator=g).bfloat16()
    a = big[:, 768:1536]  # strided view: lda = 2304
    w = torch.randn(256, 768, device=cuda, generator=g).bfloat16()
    out = ops.gemm(a, w, out_fp32=True)
    assert out.dtype == torch.float32
    _check(out, a.float() @ w.float().t(), tol=1e-3)


@pytest.mark.parametrize("M,N,K,split", [(512, 768, 50304, 12), (256, 256, 640, 4), (100, 64, 4096, 7)])
def test_gemm_split_k(cuda, M, N, K, split):
    """Few output tiles + huge contraction (d h = d logits . W): K is cut into slices whose fp32 partial tiles are
    summed in a fixed order by vlk_gemm_bf16_splitk (deterministic: two runs agree bit for bit)."""
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = torch.randn(K, N, device=cuda, generator=g).bfloat16()
    out = ops.gemm(a, b, trans_b=True, split_k=split)
    assert out.dtype == torch.bfloat16
    _check(out, a.float() @ b.float())
    assert torch.equal(out, ops.gemm(a, b, trans_b=True, split_k=split))
    with pytest.raises(RuntimeError):   # split-K is a raw-accumulate mode: no epilogue operands
        ops.gemm(a, b, trans_b=True, bias=torch.zeros(N, device=cuda).bfloat16(), split_k=split)
    acc = torch.zeros(M, N, device=cuda)
    ops.gemm(a, b, trans_b=True, out=acc, out_fp32=True, split_k=-split)          # raw atomic mode
    _check(acc, a.float() @ b.float(), tol=1e-3)
    with pytest.raises(RuntimeError):
        ops.gemm(a, b, trans_b=True, bias=torch.zeros(N, device=cuda).bfloat16(), out_fp32=True, split_k=-split)


@pytest.mark.parametrize("rows,n_out,k_in", [(16384, 768, 768), (16384, 2304, 768), (4096, 768, 3072), (2112, 768, 768),
                                             (640, 256, 128)])
def test_wgrad_auto_split_k(cuda, rows, n_out, k_in):
    """dW = dy^T . x (both operands MN-major) with the automatically chosen number of contraction slices."""
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + n_out)
    dy = (torch.randn(rows, n_out, device=cuda, generator=g) * 0.1).bfloat16()
    x = torch.randn(rows, k_in, device=cuda, generator=g).bfloat16()
    split = ops.auto_split_k(n_out, k_in, rows)
    if rows >= 16384 and n_out * k_in <= 2304 * 768:
        assert split > 1
    dw = ops.wgrad(dy, x)
    assert dw.shape == (n_out, k_in)
    _check(dw, dy.float().t() @ x.float())


@pytest.mark.parametrize("M,N,K,act", [(16448, 3072, 1024, None), (2056, 4096, 1024, "quick_gelu"), (514, 768, 1024, None),
                                        (300, 128, 64, "gelu_tanh"), (77, 256, 2048, None), (514, 64, 128, None),
                                        (40, 40, 64, "quick_gelu")])
def test_gemm_with_folded_layernorm(cuda, M, N, K, act):
    """act(LayerNorm(x) W^T + b) with the norm folded into the weights and applied in the GEMM epilogue from
    per-row statistics, vs torch fp32 on the same bf16 inputs (rows with a large common offset included: the
    epilogue subtracts mean * colsum from the raw product)."""
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device=cuda, generator=g)
    x[::7] += 6.0                                            # large-mean rows
    x[:, ::97] *= 12.0                                       # outlier channels
    x = x.bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) * 0.03).bfloat16()
    b = (torch.randn(N, device=cuda, generator=g) * 0.1).bfloat16()
    gamma = (1 + 0.2 * torch.randn(K, device=cuda, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(K, device=cuda, generator=g)).bfloat16()
    mean, rstd = ops.row_stats(x, 1e-5)
    xf = x.float()
    assert torch.allclose(mean, xf.mean(-1), atol=1e-3, rtol=1e-4)
    assert torch.allclose(rstd, (xf.var(-1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-3)
    wf, colsum, biasf = ops.fold_layernorm(w, b, gamma, beta)
    out = ops.gemm_lnfold(x, wf, biasf, colsum, 1e-5, act=act)
    ref = torch.nn.functional.layer_norm(xf, (K,), gamma.float(), beta.float(), 1e-5) @ w.float().t() + b.float()
    if act == "quick_gelu":
        ref = ref * torch.sigmoid(1.702 * ref)
    elif act == "gelu_tanh":
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    _check(out, ref, tol=2e-2)
    # and against the unfused pair on the same kernels
    h, _, _ = ops.layernorm_fwd(x, gamma, beta, 1e-5, save_stats=False)
    pair = ops.gemm(h, w, bias=b, act=act)
    assert (out.float() - ref).abs().mean().item() <= 1.5 * (pair.float() - ref).abs().mean().item() + 1e-4


@pytest.mark.parametrize("M,N,K", [(16448, 1024, 1024), (2056, 1024, 4096), (300, 128, 64), (77, 264, 72)])
def test_residual_gemm_produces_row_statistics(cuda, M, N, K):
    """vlk_gemm_bf16_stats: same output as the plain residual GEMM, plus (sum, sum of squares) per row of the ROUNDED
    output; and a folded-LayerNorm product fed with those sums agrees with the one fed by vlk_row_stats."""
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    w = (torch.randn(N, K, device=cuda, generator=g) * 0.05).bfloat16()
    b = (torch.randn(N, device=cuda, generator=g) * 0.1).bfloat16()
    res = torch.randn(M, N, device=cuda, generator=g)
    res[::5] += 4.0
    res = res.bfloat16()
    sums = torch.zeros(M, 2, device=cuda)
    out = ops.gemm_stats(a, w, b, res, sums)
    assert torch.equal(out, ops.gemm(a, w, bias=b, residual=res))
    of = out.float()
    assert torch.allclose(sums[:, 0], of.sum(-1), rtol=1e-4, atol=2e-2)
    assert torch.allclose(sums[:, 1], (of * of).sum(-1), rtol=1e-4, atol=2e-2)
    # consumer: LN folded into a following Linear, statistics from the sums vs from a pass over `out`
    N2 = 256
    w2 = (torch.randn(N2, N, device=cuda, generator=g) * 0.03).bfloat16()
    gamma = (1 + 0.2 * torch.randn(N, device=cuda, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(N, device=cuda, generator=g)).bfloat16()
    wf, colsum, biasf = ops.fold_layernorm(w2, None, gamma, beta)
    y_sums = ops.gemm_lnfold(out, wf, biasf, colsum, 1e-5, act="quick_gelu", sums=sums)
    y_stats = ops.gemm_lnfold(out, wf, biasf, colsum, 1e-5, act="quick_gelu")
    ref = torch.nn.functional.layer_norm(of, (N,), gamma.float(), beta.float(), 1e-5) @ w2.float().t()
    ref = ref * torch.sigmoid(1.702 * ref)
    _check(y_sums, ref, tol=2e-2)
    assert (y_sums.float() - y_stats.float()).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_gemm_rejects_bad_args(cuda):
    from gpt2_vision_language_b200 import ops
    a = torch.randn(16, 12, device=cuda).bfloat16()
    w = torch.randn(16, 12, device=cuda).bfloat16()
    with pytest.raises(RuntimeError):
        ops.gemm(a, w)  # K % 8 != 0
