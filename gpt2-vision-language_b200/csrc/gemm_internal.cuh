// Library-internal face of the tcgen05 GEMM (gemm_tcgen05.cu) for the other translation units that compose
// products with it (lmhead_ce.cu).  Not part of the C ABI.
#pragma once
#include "common.cuh"

namespace vlk {

// Fused lm_head + softmax cross-entropy epilogues (see EpiParams in gemm_tcgen05.cu).
struct CeEpilogue {
    int mode;                  // 1 = forward statistics (no store to D), 2 = backward d-logits chunk
    const long long* labels;   // [M] int64, GLOBAL vocabulary ids (-100 = ignored)
    float* partial;            // mode 1: float2 [slices][M] = (running max in log2 units, sum 2^(x - max)), slice = tile_n / 2 columns
    float* label_logit;        // mode 1: [M]
    const float* lse;          // mode 2: [M] natural-log sum exp of the full row
    const float* row_scale;    // mode 2: [M]
    int col0;                  // mode 2: global vocabulary id of this chunk's column 0
};

// Output columns per CTA tile the GEMM picks for a product with N columns (each epilogue warp owns half of them).
int gemm_tile_n(int N);

int gemm_impl(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd, int transA, int transB,
              const void* bias, const void* residual, int ldr, const void* aux_in, void* aux_out, int ld_aux,
              const float* scale, int act, int dact, float alpha, int out_fp32, int split_k, long long split_stride,
              int* split_used, void* stream, const float* ln_mean = nullptr, const float* ln_rstd = nullptr,
              const float* ln_colsum = nullptr, const float* ln_sums = nullptr, float ln_eps = 0.f,
              float* stats_out = nullptr, int bn_override = 0, int pair_override = -1, const CeEpilogue* ce = nullptr);

// fp32 slabs -> bf16 (deterministic split-K second pass); out = (accumulate ? out : 0) + sum of the slabs
int splitk_reduce(const float* ws, int splits, long long slab, void* out, int M, int N, int ldd, int accumulate,
                  cudaStream_t stream);

}  // namespace vlk
