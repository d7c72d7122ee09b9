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
template <int CHUNKS, bool PARAM_GRADS, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, (PARAM_GRADS && CHUNKS <= 3) ? 2 : 1)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, bf16* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int cols, int dx_accum,
                     int grad_copies) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float g[CHUNKS][8];
    float dg[CHUNKS][8], db[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = (c * 32 + lane) * 8;
        if (col < cols) unpack8(ldg16(gamma + col), g[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dg[c][i] = 0.f;
            db[c][i] = 0.f;
        }
    }
    for (int row = blockIdx.x * kWarps + warp; row < rows; row += gridDim.x * kWarps) {
        const size_t off = static_cast<size_t>(row) * cols;
        const float mu = mean[row], rs = rstd[row];
        float dyv[CHUNKS][8], xh[CHUNKS][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * 8;
            if (col < cols) {
                unpack8(ldg16(dy + off + col), dyv[c]);
                unpack8(ldg16(x + off + col), xh[c]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xh[c][i] = (xh[c][i] - mu) * rs;
                    if (PARAM_GRADS) {
                        dg[c][i] += dyv[c][i] * xh[c][i];
                        db[c][i] += dyv[c][i];
                    }
                    dyv[c][i] *= g[c][i];
                    s1 += dyv[c][i];
                    s2 += dyv[c][i] * xh[c][i];
                }
            }
        }
        s1 = warp_sum(s1) / cols;
        s2 = warp_sum(s2) / cols;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * 8;
            if (col < cols) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = rs * (dyv[c][i] - s1 - xh[c][i] * s2);
                if (dx_accum) {
                    float p[8];
                    unpack8(*reinterpret_cast<const uint4*>(dx + off + col), p);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] += p[i];
                }
                stg16(dx + off + col, pack8(o));
            }
        }
    }
    if (PARAM_GRADS) {
        // block-level reduction over the warps, then one atomic per column per block
        __shared__ float red[kWarps][32 * 8 + 1];
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int col = (c * 32 + lane) * 8;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                __syncthreads();
#pragma unroll
                for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = pass == 0 ? dg[c][i] : db[c][i];
                __syncthreads();
                if (warp == 0 && col < cols) {
                    // `grad_copies` replicas of the accumulators spread the same-address atomics of the ~300 blocks
                    float* dst = (pass == 0 ? dgamma : dbeta) + static_cast<size_t>(blockIdx.x % grad_copies) * cols;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float s = 0.f;
#pragma unroll
                        for (int w = 0; w < kWarps; ++w) s += red[w][lane * 8 + i];
                        atomicAdd(dst + col + i, s);
                    }
                }
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

extern "C" int vlk_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean,
                                 const float* rstd, void* dx, float* dgamma, float* dbeta, int rows, int cols,
                                 int dx_accum, int grad_copies, void* stream) {
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
    const int cap = dgamma ? sms * 2 : sms * 8;  // parameter grads: few, fat blocks (one atomic per column per block)
    if (blocks > cap) blocks = cap;
    const dim3 grid(blocks), block(warps * 32);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = (cols + 255) / 256;
#define LAUNCH(C, PG, W)                                                                                        \
    layernorm_bwd_kernel<C, PG, W><<<grid, block, 0, s>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(x), \
                                                          static_cast<const bf16*>(gamma), mean, rstd,          \
                                                          static_cast<bf16*>(dx), dgamma, dbeta, rows, cols, dx_accum, \
                                                          grad_copies)
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
