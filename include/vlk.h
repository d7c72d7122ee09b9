/*
 * libvlk — C ABI of the B200 (sm_100a) kernels behind the captioning / GPT-2 training step of
 * theophile-lt/gpt2-vision-language.
 *
 * The reference has no FFI of its own: every entry point below replaces one or more PyTorch library calls
 * made from the reference's nn.Module code (file:line cited per function, paths relative to the reference
 * repo root).  Conventions:
 *   - plain C types only; all pointers are DEVICE pointers unless the name says host;
 *   - bf16 tensors are `void*` to 2-byte brain-float storage, row-major, innermost dim contiguous;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing,
 *     keeps no global mutable state and never synchronises;
 *   - return 0 on success, a negative VLK_ERR_* on invalid shape / alignment / arch, a positive value is a
 *     cudaError_t from the launch.  Nothing throws across the ABI; vlk_last_error_string() has the text.
 */
#ifndef VLK_H_
#define VLK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLK_VERSION 100 /* 0.1.0 */

#define VLK_OK 0
#define VLK_ERR_INVALID_ARG (-1)
#define VLK_ERR_ALIGNMENT (-2)
#define VLK_ERR_UNSUPPORTED (-3)
#define VLK_ERR_ARCH (-4)
#define VLK_ERR_DRIVER (-5)

/* Activation selector for GEMM epilogues. Three GELUs live on the path:
 * tanh (source/gpt2/train_gpt2.py:51), erf (source/gpt2_q_former/model.py:128),
 * quick x*sigmoid(1.702x) (HF transformers activations.py, CLIP MLP). */
#define VLK_ACT_NONE 0
#define VLK_ACT_GELU_TANH 1
#define VLK_ACT_GELU_ERF 2
#define VLK_ACT_QUICK_GELU 3

int vlk_version(void);
const char* vlk_last_error_string(void);
/* Number of SMs of the current device (148 on B200); negative on error. */
int vlk_num_sms(void);
/* Number of libvlk kernel launches issued by this process so far (bench.py reports it as gpu_launches). */
long long vlk_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
 *   D[M,N] = epi( alpha * op(A)[M,K] . op(B)[K,N] )
 *   transA == 0: A stored [M,K] (lda = row stride in elements); transA == 1: A stored [K,M].
 *   transB == 0: B stored [N,K] — the nn.Linear weight layout, i.e. x @ W^T; transB == 1: B stored [K,N].
 *   epi, in this order: v = alpha*acc; v += bias[n]; aux_out[m,n] = v; v = act(v) (dact == 0) or
 *   v *= act'(aux_in[m,n]) (dact != 0); v *= *scale; v += residual[m,n]; D[m,n] = bf16(v)
 *   (or fp32 when out_fp32 != 0).  Any of bias / residual / aux_in / aux_out / scale may be NULL.
 * Replaces: nn.Linear / F.linear at train_gpt2.py:35,42,56-58,121; gpt2_linear/model.py:128;
 *   gpt2_cross-att/model.py:49-57,83; gpt2_q_former/model.py:126-130,160 and the in/out projections of
 *   nn.MultiheadAttention (:119,:123); HF modeling_clip.py q/k/v/out_proj, fc1/fc2, patch_embedding,
 *   visual_projection; plus their autograd dgrad / wgrad GEMMs.
 *   split_k > 1 cuts the contraction into slices computed by different CTAs and added atomically into an fp32 D
 *   that the caller has zeroed (no epilogue operands allowed): for few-tile / huge-K products such as
 *   d h = d logits . W (K = 50304).
 * Requirements: M,N,K > 0; N % 8 == 0 (and M % 8 == 0 when transA); lda/ldb/ldd/ldr/ld_aux % 8 == 0;
 *   16-byte aligned bases.  K is free (the tail of the last 64-wide k-block is zero-filled by TMA).
 */
int vlk_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                  int transA, int transB, const void* bias, const void* residual, int ldr, const void* aux_in,
                  void* aux_out, int ld_aux, const float* scale, int act, int dact, float alpha, int out_fp32,
                  int split_k, void* stream);

/* The same product with an explicit tile shape instead of the built-in choice (tile_n: 0 = automatic, 64 / 128 / 256
 * output columns per CTA tile; cta_pair: -1 = automatic, 0 = one CTA per 128-row tile, 1 = cta_group::2 pair on a
 * 256-row tile).  Every tile shape computes the same result; this entry exists for tuning sweeps and for the
 * regression tests that exercise each kernel instantiation.  bias / residual / act as in vlk_gemm_bf16. */
int vlk_gemm_bf16_tile(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                       int transA, int transB, const void* bias, const void* residual, int ldr, int act, int tile_n,
                       int cta_pair, void* stream);

/* Deterministic split-K product for few-tile / long-K shapes — the weight gradients dW = dy^T . x of nn.Linear
 * (autograd of train_gpt2.py:35,42,56-58: [768..3072] x [768..3072] outputs contracted over B*T = 16,384 rows fill
 * only 9..36 of the 74 tile slots) and d h = d logits . W (K = 50,304):
 *   slice s of the contraction writes its fp32 partial tile into workspace[s][M][N] (plain vector stores, no
 *   atomics); a second kernel sums the slices in a fixed order and writes D = bf16(alpha * A.B) — or
 *   D = bf16(D + alpha * A.B) when accumulate != 0 (gradient accumulation over micro-batches,
 *   train_gpt2.py:458-469).  workspace: split_k * M * N floats, 16-byte aligned, owned by the caller. */
int vlk_gemm_bf16_splitk(const void* A, const void* B, void* D, float* workspace, int M, int N, int K, int lda,
                         int ldb, int ldd, int transA, int transB, float alpha, int split_k, int accumulate,
                         void* stream);

/* LayerNorm folded into the following Linear, for FROZEN weights (the CLIP tower: layer_norm1 -> q/k/v,
 * layer_norm2 -> fc1, post_layernorm -> visual_projection; modeling_clip.py:347-365, 1026):
 *     LN(x) W^T + b  =  rstd_m * ( x Wf^T  -  mean_m * colsum_n )  +  biasf_n
 * with Wf = W * gamma (column-scaled, precomputed once), colsum_n = sum_k Wf[n,k] (fp32), biasf = W beta + b.
 * The GEMM runs on the RAW activations; the per-row statistics come from vlk_row_stats (one read of x, no
 * normalised copy is ever written) and are applied in the epilogue together with bias and activation
 * (act: VLK_ACT_NONE / QUICK_GELU / GELU_TANH).  X [M,K], Wf [N,K], D [M,N] bf16. */
int vlk_row_stats(const void* x, float* mean, float* rstd, int rows, int cols, float eps, void* stream);
int vlk_gemm_bf16_lnfold(const void* X, const void* Wf, void* D, int M, int N, int K, int ldx, int ldw, int ldd,
                         const void* bias, const float* row_mean, const float* row_rstd, const float* col_sum,
                         int act, void* stream);
/* The same with the statistics supplied as per-row sums: row_sums[m] = (sum_k x[m,k], sum_k x[m,k]^2), fp32 [M][2],
 * as accumulated by the GEMM that PRODUCED x (vlk_gemm_bf16_stats): mean = s1/K, rstd = rsqrt(s2/K - mean^2 + eps).
 * With both, a frozen pre-LN transformer block (CLIP: x += out_proj(..); fc1(LN(x))) needs no pass over x between
 * the residual GEMM and the next product at all. */
int vlk_gemm_bf16_lnfold_sums(const void* X, const void* Wf, void* D, int M, int N, int K, int ldx, int ldw, int ldd,
                              const void* bias, const float* row_sums, float eps, const float* col_sum, int act,
                              void* stream);
/* D = bf16(A.B^T + bias + residual) and, in the same epilogue, stats_out[m] += (sum_n D[m,n], sum_n D[m,n]^2) over the
 * bf16-rounded outputs (fp32 atomics into a caller-zeroed [M][2] buffer).  A [M,K], B [N,K]; N >= 96. */
int vlk_gemm_bf16_stats(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                        const void* bias, const void* residual, int ldr, float* stats_out, void* stream);

/* out[n] (fp32, overwritten) = sum_m X[m,n]; used for bias gradients (autograd of nn.Linear). */
int vlk_colsum_bf16(const void* X, float* out, int rows, int cols, int ldx, void* stream);
/* grad[n] (bf16) += sum_m X[m,n] in ONE launch (autograd's AccumulateGrad of nn.Linear.bias, train_gpt2.py:458-469):
 * `workspace` (>= cols fp32) and `counters` (>= ceil(cols / 256) uint32) must be zero on entry and are zero again on
 * exit — the last block of every 256-column group folds the sums into `grad`. */
int vlk_colsum_bf16_acc(const void* X, float* workspace, unsigned int* counters, void* grad, int rows, int cols, int ldx,
                        void* stream);

/* dst[c, r] = src[r, c] for bf16 matrices (ld in elements). */
int vlk_transpose_bf16(const void* src, void* dst, int rows, int cols, int ld_src, int ld_dst, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim, fp32 statistics (nn.LayerNorm at train_gpt2.py:66-72,94;
 * gpt2_cross-att/model.py:91-95; gpt2_q_former/model.py:118-125; CLIP pre/post/layer norms).
 * fwd: y = (x-mean)*rstd*gamma + beta; mean/rstd (fp32 [rows]) may be NULL for inference.
 * bwd: dx always; dgamma/dbeta (fp32 [grad_copies, cols], ACCUMULATED into; block b adds to replica
 *      b % grad_copies so that the ~300 blocks do not serialise on 2 x cols addresses — sum the replicas with
 *      vlk_sum_copies) may both be NULL for frozen norms.  If dx_accum != 0, dx += result (residual-stream
 *      accumulation).
 */
int vlk_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                      int rows, int cols, float eps, void* stream);
int vlk_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                      void* dx, float* dgamma, float* dbeta, int rows, int cols, int dx_accum, int grad_copies,
                      void* stream);
/* vlk_layernorm_bwd with the parameter gradients ADDED into the parameters' bf16 gradients (autograd's
 * AccumulateGrad for nn.LayerNorm.weight / .bias, train_gpt2.py:458-469) by the same launch: the blocks accumulate into
 * `grad_copies` zeroed fp32 replicas dgamma_ws / dbeta_ws [grad_copies][cols]; the block that finishes last (counter,
 * one zeroed uint32) folds them into dgamma_grad / dbeta_grad and leaves replicas and counter zeroed again. */
int vlk_layernorm_bwd_acc(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                          void* dx, float* dgamma_ws, float* dbeta_ws, int rows, int cols, int dx_accum, int grad_copies,
                          void* dgamma_grad, void* dbeta_grad, unsigned int* counter, void* stream);
/* dst[i] = (accumulate ? dst[i] : 0) + sum_c src[c*n + i], written as fp32 (dst_bf16 == 0) or bf16.  accumulate adds
 * into an existing gradient (p.grad += g of autograd's AccumulateGrad, without the extra kernel); clear_src zeroes
 * the replicas after reading them, so a persistent accumulator workspace is clean for its next user. */
int vlk_sum_copies(float* src, int copies, long long n, void* dst, int dst_bf16, int accumulate, int clear_src,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Scaled-dot-product attention, head dim 64, fp32 softmax (F.scaled_dot_product_attention at
 * train_gpt2.py:40, gpt2_cross-att/model.py:55; nn.MultiheadAttention math at gpt2_q_former/model.py:135,140;
 * CLIPAttention).  Q/K/V/O are addressed as  base + b*batch_stride + t*row_stride + h*64  (elements), so the
 * packed c_attn / kv_proj / in_proj outputs are consumed in place with no head transpose.
 * lse (fp32 [B,H,Tq], natural log) is written when non-NULL and is required by the backward.
 * bwd computes dQ, dK, dV (same addressing as their primals; written, not accumulated).
 * dropout_p > 0 drops attention probabilities (nn.MultiheadAttention(dropout=0.1), gpt2_q_former/model.py:119,123);
 * implemented for Tq, Tk <= 64 only.  Masks come from Philox-4x32-10 keyed by seed_state = device {seed, step}
 * and the call-site stream_id; the backward regenerates the forward's mask from the same triple.
 */
int vlk_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                 long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                 int o_rs, int causal, float scale, float dropout_p, const unsigned long long* seed_state,
                 unsigned int stream_id, void* stream);
int vlk_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                 void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs,
                 long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs,
                 int dq_rs, long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale,
                 float* delta_scratch /* fp32 [B*H*Tq], overwritten */, float dropout_p,
                 const unsigned long long* seed_state, unsigned int stream_id, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * 257 -> 33 token pooling: keep CLS, average the 16x16 patch grid into 4 rows x 8 cols of bins
 * (bin (r,c) = patch rows 4r..4r+3, cols 2c..2c+1), then L2-normalise every token (eps 1e-12).
 * pool_clip_197_to_33_avg_with_cls at gpt2_linear/model.py:240-254 (same in the other two model.py).
 * in: [B,257,D] (bf16 or fp32 by in_fp32), out: [B,33,D] same dtype.  normalize == 0 skips the L2 step.
 */
int vlk_pool33_l2norm(const void* in, void* out, int B, int D, int in_fp32, int normalize, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * out[b, prefix_len + t, :] = wte[ids[b,t], :] + wpe[pos0 + t, :];  out[b, 0:prefix_len, :] = prefix[b]
 * (prefix may be NULL with prefix_len == 0).  GPT.forward train_gpt2.py:114-117 and the caption
 * variant gpt2_linear/model.py:187-200 (image prefix gets no position embedding).
 * ids are int64 as produced by the reference data path; an id outside [0, vocab) poisons its output row with NaN
 * (torch's embedding raises a device assert there) instead of reading outside wte.
 */
int vlk_embed_concat_fwd(const long long* ids, const void* wte, const void* wpe, const void* prefix, void* out,
                         int B, int T, int prefix_len, int C, int pos0, int vocab, void* stream);
/* Gradient of the above w.r.t. wte / wpe (fp32 accumulators [V,C], [block,C]); pretraining only. */
int vlk_embed_bwd(const long long* ids, const void* dout, float* dwte, float* dwpe, int B, int T, int prefix_len,
                  int C, void* stream);
/* The same, ACCUMULATED into bf16 gradients (the flat bucket; wte.grad also receives the tied lm_head's dW) without
 * dense [V, C] temporaries: rows of dout that share a token id are summed in fp32 in scratch[first position of that
 * id] (fp32 [B*T, C]), then every first position adds its row into dwte[id] once.  first_pos (int32 [V]) and scratch
 * are caller-owned workspaces that must hold INT32_MAX / zeros on entry and are restored on exit.
 * dwpe[t] += sum_b dout[b, prefix_len + t].  Either gradient pointer may be NULL. */
int vlk_embed_bwd_acc(const long long* ids, const void* dout, void* dwte, void* dwpe, int* first_pos, float* scratch,
                      int B, int T, int prefix_len, int C, int vocab, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Row-wise softmax cross-entropy statistics over materialised bf16 logits [rows, V] (ld in elements):
 *   loss_row[r] = logsumexp(logits[r]) - logits[r, label[r]]   (0 for ignored rows: label == -100, the
 *   ignore_index of F.cross_entropy; NaN for any other label outside [0, V) — torch asserts there)
 * and, in place, logits[r,:] <- (softmax(logits[r]) - onehot(label[r])) * grad_scale[r]  when
 * write_grad != 0, where the caller supplies *inv_count (device fp32 scalar = 1/number of valid rows) and
 * an optional per-row fp32 weight (x-attn masked mean, gpt2_cross-att/model.py:176-185).
 * F.cross_entropy at train_gpt2.py:124; gpt2_linear/model.py:205-210 (ignore_index=-100).
 * The vocab is processed chunk-by-chunk by the caller (lm_head GEMM per row block) so that the [M,50304]
 * logits are never resident at once; see lmhead_ce in the host package.
 */
int vlk_softmax_ce_rows(void* logits, const long long* labels, const float* row_weight, float* loss_row,
                        const float* inv_count, int rows, int V, int ld, int write_grad, void* stream);
/* ---------------------------------------------------------------------------------------------------
 * Fused lm_head + softmax cross-entropy WITHOUT a [rows, V] logits buffer, forward and backward.
 * Replaces  logits = lm_head(x); F.cross_entropy(logits.view(-1, V), targets.view(-1))  and its autograd at
 * train_gpt2.py:121-124, gpt2_linear/model.py:172,204-210 (ignore_index = -100) and the masked mean of
 * gpt2_cross-att/model.py:176-185 (row_weight = mask).
 *   h [rows, C] bf16 (row stride ldh), W [V, C] bf16 (the tied wte / lm_head weight, row stride ldw), labels int64
 *   [rows] (-100 = ignored; any other value outside [0, V) makes that row's loss NaN), row_weight fp32 [rows] or NULL.
 * fwd: one tcgen05 GEMM whose epilogue reduces every 128-column slice of the logit tile to a running (max, sum exp)
 *   per row and picks the label logit — nothing of size rows x V is written —, then a merge kernel:
 *     lse[r] = logsumexp(h[r] . W^T),  loss_row[r] = lse[r] - logit[r, label[r]]  (0 if ignored),
 *     loss_out[0] = sum_r loss_row[r] * w[r] / max(count, 1),  loss_out[1] = 1 / max(count, 1),
 *     count = number of non-ignored rows, or sum of row_weight when given.
 * bwd: walks the vocabulary in chunks whose d-logits [rows, Vc] stay L2-resident; per chunk one GEMM recomputes the
 *   logit tiles and writes  (exp(logit - lse[r]) - [v == label[r]]) * w[r] * inv_count * dloss  straight from its
 *   epilogue, then  dh (+)= that . W[chunk]  (deterministic split-K)  and, when dW != NULL (trainable head:
 *   pretraining),  dW[chunk] = that^T . h  — accumulated into dW when dw_accumulate != 0 (gradient accumulation,
 *   train_gpt2.py:458-469).  inv_count = &loss_out[1] of the forward; dloss: device fp32 scalar (upstream gradient,
 *   e.g. 1/grad_accum) or NULL for 1.  dh [rows, C] bf16 is overwritten (may be NULL when only dW is wanted).
 * row_block / chunk_cols (backward geometry; 0 = automatic): rows per pass over the vocabulary and vocabulary columns
 *   per chunk.  Automatic = 4,096 rows x the widest chunk whose d-logits fit a 48 MB L2 budget (never a full
 *   [rows, V] buffer).  DESIGN.md gives the measured cost of that choice against wider chunks.
 * workspace: vlk_lmhead_ce_workspace_bytes(rows, C, V, backward, row_block, chunk_cols) bytes, 16-byte aligned, caller-owned; contents are
 *   scratch (nothing is carried from fwd to bwd except lse and loss_out[1]).
 */
long long vlk_lmhead_ce_workspace_bytes(int rows, int C, int V, int backward, int row_block, int chunk_cols);
int vlk_lmhead_ce_fwd(const void* h, const void* W, const long long* labels, const float* row_weight, float* loss_out,
                      float* loss_row, float* lse, int rows, int C, int V, int ldh, int ldw, void* workspace,
                      long long workspace_bytes, void* stream);
int vlk_lmhead_ce_bwd(const void* h, const void* W, const long long* labels, const float* row_weight, const float* lse,
                      const float* inv_count, const float* dloss, void* dh, void* dW, int dw_accumulate, int rows, int C,
                      int V, int ldh, int ldw, int lddh, int lddw, int row_block, int chunk_cols, void* workspace,
                      long long workspace_bytes, void* stream);

/* valid-count + mean: out[0] = sum(loss_row*w)/max(count,1), out[1] = 1/max(count,1), count = #labels != -100
 * (or sum of weights when row_weight != NULL). */
int vlk_ce_count(const long long* labels, const float* row_weight, float* out, int rows, void* stream);
int vlk_ce_finalize(const float* loss_row, const float* row_weight, float* out, int rows, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fused global-norm clip + AdamW over a table of tensors (clip_grad_norm_ at train_gpt2.py:472 and
 * torch.optim.AdamW(betas, eps, fused) at train_gpt2.py:143).
 *   pass 1  vlk_grad_sumsq: norm_sq[0] += sum over all grads of g^2  (fp32, caller zeroes; deterministic)
 *   pass 2  vlk_adamw_step: clip = min(1, max_norm/(sqrt(norm_sq)+1e-6)); g *= clip; standard decoupled AdamW.
 * The table is a device array of vlk_tensor_desc; params/grads/moments are bf16 or fp32 per `dtype_fp32`.
 * `grads` are left scaled?  No: gradients are read-only here; the clip factor is applied in registers.
 */
typedef struct vlk_tensor_desc {
    void* param;
    const void* grad;
    void* exp_avg;
    void* exp_avg_sq;
    long long numel;
    float weight_decay;
    int pad_;
} vlk_tensor_desc;

/* partials: caller-owned scratch of vlk_grad_sumsq_workspace_floats(n_tensors, max_numel) floats — block partial sums,
 * added in a fixed order so that every rank of a data-parallel job computes bit-identical norms from identical
 * (all-reduced) gradients. */
long long vlk_grad_sumsq_workspace_floats(int n_tensors, long long max_numel);
int vlk_grad_sumsq(const vlk_tensor_desc* table, int n_tensors, long long max_numel, int dtype_fp32,
                   float* norm_sq, float* partials, void* stream);
int vlk_adamw_step(const vlk_tensor_desc* table, int n_tensors, long long max_numel, int dtype_fp32,
                   const float* norm_sq, float max_norm, const float* lr, float beta1, float beta2, float eps,
                   const float* step, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Elementwise helpers on the path.
 */
/* CLIP patch embedding as a GEMM: unfold pixel_values [B,3,224,224] (fp32 or bf16) into
 * [B*256, Kpad] bf16 rows (k = c*196 + i*14 + j, zero-padded to Kpad). modeling_clip.py:148-154,209. */
int vlk_im2col_patch14(const void* pixels, void* out, int B, int Kpad, int in_fp32, void* stream);
/* x[b, 0, :] = cls + pos[0]; x[b, 1+p, :] = patch[b*256+p, :] + pos[1+p]   (modeling_clip.py:209-218) */
int vlk_clip_assemble(const void* patch, const void* cls, const void* pos, void* out, int B, int D, void* stream);
/* y = a + b (bf16), n elements, n % 8 == 0 */
int vlk_add_bf16(const void* a, const void* b, void* y, long long n, void* stream);
/* dst(bf16) <- src(fp32) and back */
int vlk_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
int vlk_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream);
/* y = residual + x * keep / (1 - p) with the Philox mask of (seed_state, stream_id); residual may be NULL.
 * nn.Dropout on the three residual branches of QFormerLayer (gpt2_q_former/model.py:131,136,141,144); the
 * backward is the same call on dy with residual = NULL. n % 8 == 0. */
int vlk_dropout_add_bf16(const void* x, const void* residual, void* y, long long n, float p,
                         const unsigned long long* seed_state, unsigned int stream_id, void* stream);
/* gate gradient for the x-attn block (gpt2_cross-att/model.py:101):
 * out[0] += (1 - tanh(gate)^2) * sum(dy * y)  over n elements. */
int vlk_gate_grad(const void* dy, const void* y, const float* gate, float* out, long long n, void* stream);
/* A scalar riding in the gradient all-reduce (the averaged loss of train_gpt2.py:470-471 without a collective of its
 * own): value (device fp32, 0 <= v < 256) <-> 8 slots of the flat gradient bucket (bf16, or fp32 when slots_fp32)
 * holding the base-16 digits of its Q8.24 fixed-point form.  Digit sums over <= 16 ranks and the division by a
 * power-of-two world size are exact, so unpack(average(pack(v_r))) = mean(v_r) to 2^-24. */
int vlk_scalar_pack_digits(const float* value, void* slots, int slots_fp32, void* stream);
int vlk_scalar_unpack_digits(const void* slots, float* value, int slots_fp32, void* stream);
/* argmax over the last dim of bf16 [rows, V] -> int64 ids (greedy decode). */
int vlk_argmax_rows(const void* logits, long long* out, int rows, int V, int ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLK_H_ */
