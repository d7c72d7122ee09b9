// Fused lm_head + softmax cross-entropy, forward and backward, without a [rows, V] logits buffer.
//
// Replaces  logits = lm_head(x); loss = F.cross_entropy(logits.view(-1, V), targets.view(-1))  and its autograd
// (source/gpt2/train_gpt2.py:121-124; gpt2_linear/model.py:172,204-210 with ignore_index = -100; the masked mean of
// gpt2_cross-att/model.py:176-185 through row_weight).
//
// forward   one tcgen05 GEMM  h [rows, C] x W^T [C, V]  whose epilogue never stores the logits: every epilogue warp keeps
//           a running (max, sum exp) per row over its 128 columns of the accumulator tile (TMEM) and drops one float2 per
//           (row, 128-column slice); the thread that meets the label column writes that logit.  A second, tiny kernel
//           merges the V/128 partials of a row into lse[row] and loss_row[row] = lse - logit[label].
//           Traffic: W once (77 MB at V = 50304, C = 768), h, 8 B per row and slice of partials.  No logits.
// backward  the vocabulary is walked in chunks sized so that a chunk of d-logits [rows, Vc] (bf16) stays L2-resident:
//           per chunk one GEMM RECOMPUTES the logit tiles and its epilogue writes
//               dL = (exp(logit - lse[row]) - [col == label[row]]) * row_scale[row]
//           then  dh (+)= dL . W[chunk]   (deterministic split-K, fp32 slabs, accumulated into bf16 dh)
//           and   dW[chunk] (+)= dL^T . h (only when the head is trainable: pretraining).
//           The chunk buffer is the only [rows x Vc] array that ever exists; it is rewritten every chunk.
//
// row_scale[row] = row_weight[row] * (1 / count) * dloss  (0 for ignored rows), so the upstream gradient and the
// 1/grad_accum factor of the training loop are folded in here — no separate scaling pass over dh / dW.
#include "common.cuh"
#include "gemm_internal.cuh"

namespace vlk {
namespace {

constexpr long long kChunkBytes = 48ll << 20;   // d-logits chunk budget: well inside the 126 MB L2 next to W[chunk] and h
constexpr int kRowBlock = 4096;                 // backward: rows per pass over the vocabulary (a 16 x 1024 pretraining
                                                // micro-batch = 4 passes; the 1,984 text rows of a caption step = 1)

// One WARP per row: lanes stride the V/128 slice partials (independent loads), then a shuffle tree merges the 32
// (max, sum) pairs.  (One thread per row would chain ~400 dependent L2 round trips.)
__global__ void __launch_bounds__(256)
ce_combine_kernel(const float2* __restrict__ partial, const float* __restrict__ label_logit,
                  const long long* __restrict__ labels, float* __restrict__ lse, float* __restrict__ loss_row, int rows,
                  int slices, int V) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float m = -INFINITY, s = 0.f;     // log2 units, as written by the GEMM epilogue
    for (int i = lane; i < slices; i += 32) {
        const float2 p = __ldg(partial + static_cast<size_t>(i) * rows + row);
        const float nm = fmaxf(m, p.x);
        s = s * exp2f(m - nm) + p.y * exp2f(p.x - nm);
        m = nm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
        const float nm = fmaxf(m, om);
        // a lane without any slice holds (-inf, 0): exp2f(-inf - nm) = 0 (nm is finite as soon as one lane saw data)
        s = (m == -INFINITY ? 0.f : s * exp2f(m - nm)) + (om == -INFINITY ? 0.f : os * exp2f(om - nm));
        m = nm;
    }
    if (lane != 0) return;
    const float l = (m + log2f(s)) * 0.6931471805599453f;
    lse[row] = l;
    const long long label = labels[row];
    float loss = 0.f;                                   // ignore_index
    if (label != -100) loss = (label >= 0 && label < V) ? l - label_logit[row] : nanf("");   // torch asserts on a bad label
    loss_row[row] = loss;
}

// row_scale[r] = (label ignored ? 0 : weight[r]) * inv_count * dloss
__global__ void __launch_bounds__(256)
ce_row_scale_kernel(const long long* __restrict__ labels, const float* __restrict__ row_weight,
                    const float* __restrict__ inv_count, const float* __restrict__ dloss, float* __restrict__ out,
                    int rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float g = __ldg(inv_count) * (dloss ? __ldg(dloss) : 1.0f);
    out[row] = labels[row] == -100 ? 0.f : (row_weight ? row_weight[row] : 1.0f) * g;
}

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// Slices of the contraction for a product whose output has too few 256 x 256 tiles to fill the GPU (the host-side
// heuristic of ops.auto_split_k, restated): waves of tile-pair slots x (k-blocks per slice + ~6 of fixed cost).
int pick_split_k(int M, int N, int K, int sms) {
    const int slots = sms / 2 > 0 ? sms / 2 : 1;
    const int units = ((M + 255) / 256) * ((N + 255) / 256);
    const int kb = (K + 63) / 64;
    if (units * 2 > slots || kb < 32) return 1;
    int best = 1;
    long best_cost = static_cast<long>((units + slots - 1) / slots) * (kb + 6);
    for (int s = 2; s <= 16; ++s) {
        const long cost = static_cast<long>((units * s + slots - 1) / slots) * ((kb + s - 1) / s + 6);
        if (cost < best_cost) {
            best = s;
            best_cost = cost;
        }
    }
    return best;
}

int chunk_cols(int rows, int V) {
    long long vc = kChunkBytes / (2ll * rows) / 256 * 256;
    if (vc < 256) vc = 256;
    if (vc > V) vc = V;
    return static_cast<int>(vc);
}

struct BwdPlan {
    int rb, vc, split;      // rows per block, vocabulary columns per chunk, split-K of the dh product
    size_t off_scale, off_chunk, off_slabs, total;
};
BwdPlan plan_bwd(int rows, int C, int V, int sms, int row_block, int chunk) {
    BwdPlan p;
    p.rb = row_block > 0 ? row_block : kRowBlock;
    if (p.rb > rows) p.rb = rows;
    p.vc = chunk > 0 ? (chunk + 7) / 8 * 8 : chunk_cols(p.rb, V);
    if (p.vc > V) p.vc = V;
    p.split = pick_split_k(p.rb, C, p.vc, sms);
    p.off_scale = 0;
    p.off_chunk = align256(sizeof(float) * rows);
    p.off_slabs = p.off_chunk + align256(2ull * p.rb * p.vc);
    p.total = p.off_slabs + align256(sizeof(float) * static_cast<size_t>(p.split) * p.rb * C);
    return p;
}

}  // namespace
}  // namespace vlk

using namespace vlk;

extern "C" long long vlk_lmhead_ce_workspace_bytes(int rows, int C, int V, int backward, int row_block, int chunk_cols) {
    if (rows <= 0 || C <= 0 || V <= 0) return -1;
    if (!backward) {
        const int slices = (V + gemm_tile_n(V) / 2 - 1) / (gemm_tile_n(V) / 2);
        return static_cast<long long>(align256(8ull * slices * rows) + align256(4ull * rows));
    }
    int sms = device_sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<long long>(plan_bwd(rows, C, V, sms, row_block, chunk_cols).total);
}

extern "C" int vlk_lmhead_ce_fwd(const void* h, const void* W, const long long* labels, const float* row_weight,
                                 float* loss_out, float* loss_row, float* lse, int rows, int C, int V, int ldh, int ldw,
                                 void* workspace, long long workspace_bytes, void* stream) {
    VLK_REQUIRE(h && W && labels && loss_out && loss_row && lse && workspace, VLK_ERR_INVALID_ARG,
                "vlk_lmhead_ce_fwd: null pointer");
    VLK_REQUIRE(rows > 0 && C > 0 && V > 0 && V % 8 == 0, VLK_ERR_INVALID_ARG,
                "vlk_lmhead_ce_fwd: rows=%d C=%d V=%d (V must be a multiple of 8)", rows, C, V);
    VLK_REQUIRE(workspace_bytes >= vlk_lmhead_ce_workspace_bytes(rows, C, V, 0, 0, 0) && aligned16(workspace),
                VLK_ERR_INVALID_ARG, "vlk_lmhead_ce_fwd: workspace too small (%lld bytes)", workspace_bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int slice_cols = gemm_tile_n(V) / 2;
    const int slices = (V + slice_cols - 1) / slice_cols;
    float* partial = static_cast<float*>(workspace);
    float* label_logit = reinterpret_cast<float*>(static_cast<char*>(workspace) + align256(8ull * slices * rows));
    int rc = vlk_ce_count(labels, row_weight, loss_out, rows, stream);      // loss_out[1] = 1 / max(count, 1)
    if (rc) return rc;
    CeEpilogue ce = {};
    ce.mode = 1;
    ce.labels = labels;
    ce.partial = partial;
    ce.label_logit = label_logit;
    // D is never written in this mode; the partial buffer stands in for the (required, aligned) pointer
    rc = gemm_impl(h, W, partial, rows, V, C, ldh, ldw, (V + 7) / 8 * 8, 0, 0, nullptr, nullptr, 0, nullptr, nullptr, 0,
                   nullptr, VLK_ACT_NONE, 0, 1.0f, 0, 1, 0, nullptr, stream, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr,
                   0, -1, &ce);
    if (rc) return rc;
    ce_combine_kernel<<<(rows + 7) / 8, 256, 0, s>>>(reinterpret_cast<const float2*>(partial), label_logit, labels, lse,
                                                        loss_row, rows, slices, V);
    VLK_CHECK_LAUNCH("vlk_lmhead_ce_fwd(combine)");
    return vlk_ce_finalize(loss_row, row_weight, loss_out, rows, stream);   // loss_out[0] = sum(loss_row * w) / count
}

extern "C" int vlk_lmhead_ce_bwd(const void* h, const void* W, const long long* labels, const float* row_weight,
                                 const float* lse, const float* inv_count, const float* dloss, void* dh, void* dW,
                                 int dw_accumulate, int rows, int C, int V, int ldh, int ldw, int lddh, int lddw,
                                 int row_block, int chunk_cols, void* workspace, long long workspace_bytes,
                                 void* stream) {
    VLK_REQUIRE(h && W && labels && lse && inv_count && workspace && (dh || dW), VLK_ERR_INVALID_ARG,
                "vlk_lmhead_ce_bwd: null pointer");
    VLK_REQUIRE(rows > 0 && C > 0 && V > 0 && V % 8 == 0 && C % 8 == 0, VLK_ERR_INVALID_ARG,
                "vlk_lmhead_ce_bwd: rows=%d C=%d V=%d", rows, C, V);
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_lmhead_ce_bwd: no sm_100 device");
    VLK_REQUIRE(row_block >= 0 && chunk_cols >= 0, VLK_ERR_INVALID_ARG, "vlk_lmhead_ce_bwd: row_block / chunk_cols");
    const BwdPlan p = plan_bwd(rows, C, V, sms, row_block, chunk_cols);
    VLK_REQUIRE(workspace_bytes >= static_cast<long long>(p.total) && aligned16(workspace), VLK_ERR_INVALID_ARG,
                "vlk_lmhead_ce_bwd: workspace too small (%lld < %lld bytes)", workspace_bytes,
                static_cast<long long>(p.total));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    float* row_scale = reinterpret_cast<float*>(ws + p.off_scale);
    bf16* chunk = reinterpret_cast<bf16*>(ws + p.off_chunk);
    float* slabs = reinterpret_cast<float*>(ws + p.off_slabs);
    ce_row_scale_kernel<<<(rows + 255) / 256, 256, 0, s>>>(labels, row_weight, inv_count, dloss, row_scale, rows);
    VLK_CHECK_LAUNCH("vlk_lmhead_ce_bwd(row scale)");
    const bf16* Wb = static_cast<const bf16*>(W);
    const bf16* hb = static_cast<const bf16*>(h);
    const int ldc = p.vc;                               // chunk buffer row stride (multiple of 8)
    for (int r0 = 0, bi = 0; r0 < rows; r0 += p.rb, ++bi) {
        const int nr = rows - r0 < p.rb ? rows - r0 : p.rb;
        const bf16* hr = hb + static_cast<size_t>(r0) * ldh;
        for (int v0 = 0, ci = 0; v0 < V; v0 += p.vc, ++ci) {
            const int vc = V - v0 < p.vc ? V - v0 : p.vc;
            CeEpilogue ce = {};
            ce.mode = 2;
            ce.labels = labels + r0;
            ce.lse = lse + r0;
            ce.row_scale = row_scale + r0;
            ce.col0 = v0;
            // d logits of this block of rows and vocabulary chunk, recomputed: [nr, vc] = epi(h . W[v0:v0+vc]^T)
            int rc = gemm_impl(hr, Wb + static_cast<size_t>(v0) * ldw, chunk, nr, vc, C, ldh, ldw, ldc, 0, 0, nullptr, nullptr,
                               0, nullptr, nullptr, 0, nullptr, VLK_ACT_NONE, 0, 1.0f, 0, 1, 0, nullptr, stream, nullptr,
                               nullptr, nullptr, nullptr, 0.f, nullptr, 0, -1, &ce);
            if (rc) return rc;
            if (dh) {
                // dh[r0:r0+nr] (+)= dL [nr, vc] . W[v0:v0+vc, :]   (B stored [K = vc, N = C]: transB)
                int used = 1;
                const long long slab = static_cast<long long>(nr) * C;
                rc = gemm_impl(chunk, Wb + static_cast<size_t>(v0) * ldw, slabs, nr, C, vc, ldc, ldw, C, 0, 1, nullptr,
                               nullptr, 0, nullptr, nullptr, 0, nullptr, VLK_ACT_NONE, 0, 1.0f, 1, p.split, slab, &used,
                               stream);
                if (rc) return rc;
                rc = splitk_reduce(slabs, used, slab, static_cast<bf16*>(dh) + static_cast<size_t>(r0) * lddh, nr, C, lddh,
                                   ci > 0 ? 1 : 0, s);
                if (rc) return rc;
            }
            if (dW) {
                // dW[v0:v0+vc, :] (+)= dL^T [vc, nr] . h [nr, C]  (A stored [K = nr, M = vc], B stored [K = nr, N = C])
                bf16* dWc = static_cast<bf16*>(dW) + static_cast<size_t>(v0) * lddw;
                const bool acc = dw_accumulate || bi > 0;
                rc = gemm_impl(chunk, hr, dWc, vc, C, nr, ldc, ldh, lddw, 1, 1, nullptr, acc ? dWc : nullptr, lddw, nullptr,
                               nullptr, 0, nullptr, VLK_ACT_NONE, 0, 1.0f, 0, 1, 0, nullptr, stream);
                if (rc) return rc;
            }
        }
    }
    return VLK_OK;
}
