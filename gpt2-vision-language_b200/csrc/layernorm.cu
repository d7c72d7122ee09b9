// LayerNorm forward / backward: one warp per row, 16-byte vector loads, fp32 statistics via warp shuffles.
// HBM-bound: algorithmic traffic is 2 B in + 2 B out per element (3,072 B per 768-wide row).
#include "common.cuh"

namespace vlk {
namespace {

constexpr int kMaxChunks = 8;  // 8 chunks x 32 lanes x 8 elements = 2048 columns max
constexpr int kWarpsPerBlock = 4;

template <int CHUNKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                     int cols, float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const bf16* xr = x + static_cast<size_t>(row) * cols;
    float v[CHUNKS][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < cols) {
            unpack8(ldg16(xr + col), v[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += v[c][i];
        }
    }
    const float mean = warp_sum(sum) / cols;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < cols) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = v[c][i] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    bf16* yr = y + static_cast<size_t>(row) * cols;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < cols) {
            float g[8], b[8], o[8];
            unpack8(ldg16(gamma + col), g);
            unpack8(ldg16(beta + col), b);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (v[c][i] - mean) * rstd * g[i] + b[i];
            stg16(yr + col, pack8(o));
        }
    }
}

// Row statistics only (mean, rstd): the input of a GEMM whose weights have the LayerNorm folded in
// (vlk_gemm_bf16_lnfold).  One 2-byte read per element, nothing written back but 8 bytes per row.
template <int CHUNKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
row_stats_kernel(const bf16* __restrict__ x, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                 int cols, float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const bf16* xr = x + static_cast<size_t>(row) * cols;
    float v[CHUNKS][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < cols) {
            unpack8(ldg16(xr + col), v[c]);
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += v[c][i];
        }
    }
    const float mean = warp_sum(sum) / cols;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < cols) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = v[c][i] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)).  Warps walk rows grid-stride so the optional
// dgamma / dbeta partials stay in registers until one atomicAdd per column per block.
// With PARAM_GRADS the grid is only 2 blocks of 8 warps per SM: every block ends with one atomicAdd per column, and
// the contention on those 2 x cols addresses (not the streaming loop) is what bounded the kernel when it ran with
// 8 x SMs small blocks (138 us vs 25 us for dx alone at 16384 x 768).
// Software-pipelined: the NEXT row's dy / x (and its statistics) are fetched before the current row is reduced, so a
// warp always has a row of loads in flight under its two shuffle reductions (16 warps per SM could not cover the HBM
// latency otherwise: 39 us = 0.4 of the copy bandwidth at 16384 x 768 with accumulation into dx).  To make room for the
// second row the operands stay PACKED (bf16 pairs, unpacked where they are used) — gamma as well.
__device__ __forceinline__ float2 bf2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const bf162*>(&u)); }
__device__ __forceinline__ uint32_t word(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// PARAM_GRADS: the per-lane dgamma / dbeta partial sums live in a warp-private slice of (dynamic) shared memory, not in
// registers (48 accumulators per lane on top of two packed rows spilled at 128 registers): a lane owns 8 consecutive
// floats per chunk, so the read-modify-write is two conflict-free 128-bit accesses per array and chunk.
constexpr int kBwdBlocksPerSM = 2;
template <int CHUNKS, bool PARAM_GRADS, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, (PARAM_GRADS && CHUNKS <= 3) ? kBwdBlocksPerSM : 1)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, bf16* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int cols, int dx_accum,
                     int grad_copies, bf16* __restrict__ dgamma_out, bf16* __restrict__ dbeta_out,
                     unsigned int* __restrict__ done_counter) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    __shared__ uint4 sgamma[CHUNKS * 32];   // gamma, packed, in this lane's order (registers are needed for two rows)
    extern __shared__ float4 sacc4[];       // PARAM_GRADS: [kWarps][2 (dgamma | dbeta)][CHUNKS * 256] floats
    constexpr int kAccFloats = CHUNKS * 256;
    float* acc_g = reinterpret_cast<float*>(sacc4) + static_cast<size_t>(warp) * 2 * kAccFloats;
    float* acc_b = acc_g + kAccFloats;
    for (int t = threadIdx.x; t < CHUNKS * 32; t += kWarps * 32)
        sgamma[t] = t * 8 < cols ? ldg16(gamma + t * 8) : make_uint4(0, 0, 0, 0);
    if (PARAM_GRADS) {
        for (int t = lane; t < 2 * kAccFloats / 4; t += 32) reinterpret_cast<float4*>(acc_g)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int stride = gridDim.x * kWarps;
    int row = blockIdx.x * kWarps + warp;
    uint4 dyp[CHUNKS], xp[CHUNKS];    // the current row, packed
    float mu = 0.f, rs = 0.f;
    auto fetch = [&](int r, uint4 (&a)[CHUNKS], uint4 (&b)[CHUNKS], float& m_, float& r_) {
        const size_t off = static_cast<size_t>(r) * cols;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * 8;
            if (col < cols) {
                a[c] = ldg16(dy + off + col);
                b[c] = ldg16(x + off + col);
            } else {
                a[c] = make_uint4(0, 0, 0, 0);
                b[c] = make_uint4(0, 0, 0, 0);
            }
        }
        m_ = __ldg(mean + r);
        r_ = __ldg(rstd + r);
    };
    if (row < rows) fetch(row, dyp, xp, mu, rs);
    for (; row < rows; row += stride) {
        const size_t off = static_cast<size_t>(row) * cols;
        uint4 dyn[CHUNKS], xn[CHUNKS];
        float mun = 0.f, rsn = 0.f;
        const int next = row + stride;
        if (next < rows) fetch(next, dyn, xn, mun, rsn);   // in flight under this row's arithmetic and reductions
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const uint4 gpc = sgamma[c * 32 + lane];
            float pg[8], pb[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 d = bf2(word(dyp[c], i)), xv = bf2(word(xp[c], i)), gv = bf2(word(gpc, i));
                const float h0 = (xv.x - mu) * rs, h1 = (xv.y - mu) * rs;
                if (PARAM_GRADS) {
                    pg[2 * i] = d.x * h0;
                    pg[2 * i + 1] = d.y * h1;
                    pb[2 * i] = d.x;
                    pb[2 * i + 1] = d.y;
                }
                const float t0 = d.x * gv.x, t1 = d.y * gv.y;
                s1 += t0 + t1;
                s2 = fmaf(t0, h0, fmaf(t1, h1, s2));
            }
            if (PARAM_GRADS) {
                // layout [chunk][half][lane][4 floats]: every 128-bit access of the warp is one contiguous 512-byte row
                float4* ag = reinterpret_cast<float4*>(acc_g) + c * 64 + lane;
                float4* ab = reinterpret_cast<float4*>(acc_b) + c * 64 + lane;
                float4 a0 = ag[0], a1 = ag[32], b0 = ab[0], b1 = ab[32];
                a0.x += pg[0]; a0.y += pg[1]; a0.z += pg[2]; a0.w += pg[3];
                a1.x += pg[4]; a1.y += pg[5]; a1.z += pg[6]; a1.w += pg[7];
                b0.x += pb[0]; b0.y += pb[1]; b0.z += pb[2]; b0.w += pb[3];
                b1.x += pb[4]; b1.y += pb[5]; b1.z += pb[6]; b1.w += pb[7];
                ag[0] = a0; ag[32] = a1; ab[0] = b0; ab[32] = b1;
            }
        }
        s1 = warp_sum(s1) / cols;
        s2 = warp_sum(s2) / cols;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * 8;
            if (col < cols) {
                uint4 prev = make_uint4(0, 0, 0, 0);
                if (dx_accum) prev = *reinterpret_cast<const uint4*>(dx + off + col);
                const uint4 gpc = sgamma[c * 32 + lane];
                uint4 out;
                uint32_t* ow = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 d = bf2(word(dyp[c], i)), xv = bf2(word(xp[c], i)), gv = bf2(word(gpc, i));
                    const float2 pv = bf2(word(prev, i));
                    const float h0 = (xv.x - mu) * rs, h1 = (xv.y - mu) * rs;
                    const float o0 = rs * (d.x * gv.x - s1 - h0 * s2) + pv.x;
                    const float o1 = rs * (d.y * gv.y - s1 - h1 * s2) + pv.y;
                    const bf162 h2 = __floats2bfloat162_rn(o0, o1);
                    ow[i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                stg16(dx + off + col, out);
            }
        }
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            dyp[c] = dyn[c];
            xp[c] = xn[c];
        }
        mu = mun;
        rs = rsn;
    }
    if (PARAM_GRADS) {
        // block-level reduction over the warps' slices, then one atomic per column per block; `grad_copies` replicas of
        // the accumulators spread the same-address atomics of the blocks
        __syncthreads();
        const float* all = reinterpret_cast<const float*>(sacc4);
        const size_t rep = static_cast<size_t>(blockIdx.x % grad_copies) * cols;
        for (int t = threadIdx.x; t < 2 * cols; t += kWarps * 32) {
            const int which = t >= cols, col = which ? t - cols : t;
            // column -> slot of the [chunk][half][lane][4] layout
            const int c = col >> 8, ln = (col >> 3) & 31, k = col & 7;
            const int slot = ((c * 2 + (k >> 2)) * 32 + ln) * 4 + (k & 3);
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) sum += all[(static_cast<size_t>(w) * 2 + which) * kAccFloats + slot];
            atomicAdd((which ? dbeta : dgamma) + rep + col, sum);
        }
        if (dgamma_out != nullptr) {
            // The block that finishes last folds the replicas into the parameters' bf16 gradients (+=) and leaves the
            // replicas and the counter zeroed for the next call: no separate reduction launches.
            __shared__ int is_last;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
            __syncthreads();
            if (is_last) {
                __threadfence();
                for (int t = threadIdx.x; t < 2 * cols; t += kWarps * 32) {
                    const int which = t >= cols, col = which ? t - cols : t;
                    float* src = (which ? dbeta : dgamma) + col;
                    float sum = 0.f;
                    for (int c = 0; c < grad_copies; ++c) {
                        sum += __ldcg(src + static_cast<size_t>(c) * cols);
                        src[static_cast<size_t>(c) * cols] = 0.f;
                    }
                    bf16* dst = (which ? dbeta_out : dgamma_out) + col;
                    *dst = __float2bfloat16(sum + __bfloat162float(*dst));
                }
                if (threadIdx.x == 0) *done_counter = 0u;
            }
        }
    }
}

}  // namespace
}  // namespace vlk

using namespace vlk;

extern "C" int vlk_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean,
                                 float* rstd, int rows, int cols, float eps, void* stream) {
    VLK_REQUIRE(x && gamma && beta && y, VLK_ERR_INVALID_ARG, "vlk_layernorm_fwd: null pointer");
    VLK_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kMaxChunks * 256, VLK_ERR_INVALID_ARG,
                "vlk_layernorm_fwd: rows=%d cols=%d (cols must be a multiple of 8, <= 2048)", rows, cols);
    VLK_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), VLK_ERR_ALIGNMENT,
                "vlk_layernorm_fwd: 16B alignment");
    const dim3 grid((rows + kWarpsPerBlock - 1) / kWarpsPerBlock), block(kWarpsPerBlock * 32);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = (cols + 255) / 256;
#define LAUNCH(C)                                                                                              \
    layernorm_fwd_kernel<C><<<grid, block, 0, s>>>(static_cast<const bf16*>(x), static_cast<const bf16*>(gamma), \
                                                   static_cast<const bf16*>(beta), static_cast<bf16*>(y), mean,  \
                                                   rstd, rows, cols, eps)
    if (chunks <= 3) LAUNCH(3);
    else if (chunks <= 4) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    VLK_CHECK_LAUNCH("vlk_layernorm_fwd");
    return VLK_OK;
}

extern "C" int vlk_row_stats(const void* x, float* mean, float* rstd, int rows, int cols, float eps, void* stream) {
    VLK_REQUIRE(x && mean && rstd, VLK_ERR_INVALID_ARG, "vlk_row_stats: null pointer");
    VLK_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kMaxChunks * 256, VLK_ERR_INVALID_ARG,
                "vlk_row_stats: rows=%d cols=%d (cols must be a multiple of 8, <= 2048)", rows, cols);
    VLK_REQUIRE(aligned16(x), VLK_ERR_ALIGNMENT, "vlk_row_stats: 16B alignment");
    const dim3 grid((rows + kWarpsPerBlock - 1) / kWarpsPerBlock), block(kWarpsPerBlock * 32);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = (cols + 255) / 256;
    if (chunks <= 3) row_stats_kernel<3><<<grid, block, 0, s>>>(static_cast<const bf16*>(x), mean, rstd, rows, cols, eps);
    else if (chunks <= 4) row_stats_kernel<4><<<grid, block, 0, s>>>(static_cast<const bf16*>(x), mean, rstd, rows, cols, eps);
    else row_stats_kernel<8><<<grid, block, 0, s>>>(static_cast<const bf16*>(x), mean, rstd, rows, cols, eps);
    VLK_CHECK_LAUNCH("vlk_row_stats");
    return VLK_OK;
}

namespace {
int layernorm_bwd_launch(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, void* dx,
                         float* dgamma, float* dbeta, int rows, int cols, int dx_accum, int grad_copies, void* dgamma_out,
                         void* dbeta_out, unsigned int* counter, void* stream) {
    VLK_REQUIRE(dy && x && gamma && mean && rstd && dx, VLK_ERR_INVALID_ARG, "vlk_layernorm_bwd: null pointer");
    VLK_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), VLK_ERR_INVALID_ARG,
                "vlk_layernorm_bwd: dgamma and dbeta must both be given or both be NULL");
    VLK_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kMaxChunks * 256, VLK_ERR_INVALID_ARG,
                "vlk_layernorm_bwd: rows=%d cols=%d", rows, cols);
    if (grad_copies < 1) grad_copies = 1;
    VLK_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma), VLK_ERR_ALIGNMENT,
                "vlk_layernorm_bwd: 16B alignment");
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_layernorm_bwd: no sm_100 device");
    constexpr int kWarpsPG = 8;
    const int warps = dgamma ? kWarpsPG : kWarpsPerBlock;
    int blocks = (rows + warps - 1) / warps;
    const int cap = dgamma ? sms * kBwdBlocksPerSM : sms * 8;  // parameter grads: few, fat blocks (one atomic per column per block)
    if (blocks > cap) blocks = cap;
    const dim3 grid(blocks), block(warps * 32);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = (cols + 255) / 256;
#define LAUNCH(C, PG, W)                                                                                        \
    do {                                                                                                        \
        const int dyn = PG ? W * 2 * C * 256 * 4 : 0;                                                           \
        static bool configured = false;                                                                         \
        if (PG && !configured) {                                                                                \
            VLK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<C, PG, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)); \
            configured = true;                                                                                  \
        }                                                                                                       \
        layernorm_bwd_kernel<C, PG, W><<<grid, block, dyn, s>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(x), \
                                                            static_cast<const bf16*>(gamma), mean, rstd,        \
                                                            static_cast<bf16*>(dx), dgamma, dbeta, rows, cols, dx_accum, \
                                                            grad_copies, static_cast<bf16*>(dgamma_out),        \
                                                            static_cast<bf16*>(dbeta_out), counter);            \
    } while (0)
    if (dgamma) {
        if (chunks <= 3) LAUNCH(3, true, kWarpsPG);
        else if (chunks <= 4) LAUNCH(4, true, kWarpsPG);
        else LAUNCH(8, true, kWarpsPG);
    } else {
        if (chunks <= 3) LAUNCH(3, false, kWarpsPerBlock);
        else if (chunks <= 4) LAUNCH(4, false, kWarpsPerBlock);
        else LAUNCH(8, false, kWarpsPerBlock);
    }
#undef LAUNCH
    VLK_CHECK_LAUNCH("vlk_layernorm_bwd");
    return VLK_OK;
}
}  // namespace

extern "C" int vlk_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                                 const float* rstd, void* dx, float* dgamma, float* dbeta, int rows, int cols,
                                 int dx_accum, int grad_copies, void* stream) {
    return layernorm_bwd_launch(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, cols, dx_accum, grad_copies, nullptr,
                                nullptr, nullptr, stream);
}

extern "C" int vlk_layernorm_bwd_acc(const void* dy, const void* x, const void* gamma, const float* mean,
                                     const float* rstd, void* dx, float* dgamma_ws, float* dbeta_ws, int rows, int cols,
                                     int dx_accum, int grad_copies, void* dgamma_grad, void* dbeta_grad,
                                     unsigned int* counter, void* stream) {
    VLK_REQUIRE(dgamma_ws && dbeta_ws && dgamma_grad && dbeta_grad && counter, VLK_ERR_INVALID_ARG,
                "vlk_layernorm_bwd_acc: null pointer");
    return layernorm_bwd_launch(dy, x, gamma, mean, rstd, dx, dgamma_ws, dbeta_ws, rows, cols, dx_accum, grad_copies,
                                dgamma_grad, dbeta_grad, counter, stream);
}

// dst[i] = (accumulate ? dst[i] : 0) + sum_c src[c][i]  (fp32 or bf16 destination): final reduction of replicated
// gradient accumulators, optionally ADDED to an existing gradient (the flat bucket: gradient accumulation over
// micro-batches without a separate add kernel) and optionally clearing the source so that a persistent accumulator
// workspace is zero again for its next user.
__global__ void sum_copies_kernel(float* __restrict__ src, int copies, long long n, float* __restrict__ dst_f32,
                                  __nv_bfloat16* __restrict__ dst_bf16, int accumulate, int clear_src) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < copies; ++c) {
            s += src[c * n + i];
            if (clear_src) src[c * n + i] = 0.f;
        }
        if (dst_bf16) dst_bf16[i] = __float2bfloat16(accumulate ? s + __bfloat162float(dst_bf16[i]) : s);
        else dst_f32[i] = accumulate ? s + dst_f32[i] : s;
    }
}

extern "C" int vlk_sum_copies(float* src, int copies, long long n, void* dst, int dst_bf16, int accumulate,
                              int clear_src, void* stream) {
    VLK_REQUIRE(src && dst && copies > 0 && n > 0, VLK_ERR_INVALID_ARG, "vlk_sum_copies: args");
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    sum_copies_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, copies, n, dst_bf16 ? nullptr : static_cast<float*>(dst), dst_bf16 ? static_cast<__nv_bfloat16*>(dst) : nullptr,
        accumulate, clear_src);
    VLK_CHECK_LAUNCH("vlk_sum_copies");
    return VLK_OK;
}
