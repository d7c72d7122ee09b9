// bf16 GEMM on the 5th-gen tensor cores: TMA (128B-swizzled tiles) -> shared memory ring ->
// tcgen05.mma (one issuing thread, fp32 accumulators in tensor memory, double-buffered) -> epilogue warps
// (tcgen05.ld, bias / activation / activation-gradient / scale / residual in registers, bf16 stores).
//
// Persistent: one CTA per SM walks output tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  with the M index
// fastest, so that the CTAs of one wave share the same weight tile (L2 reuse) while activations stream.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 idle, warps 4..11 = epilogue (warp w owns TMEM lane quadrant w%4 and column half (w-4)/4).
//
// Operand layouts: an operand is "K-major" when the contraction index is contiguous in global memory
// (activations [M,K], nn.Linear weights [N,K]) and "MN-major" when the M/N index is contiguous
// (x^T for weight gradients, W[N_out,K_in] read as [K=N_out, N=K_in] for input gradients).  Both are fed
// to tcgen05.mma directly through the descriptor major bits; no transposes are materialised.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vlk {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;

struct EpiParams {
    const bf16* bias;
    const bf16* residual;
    const bf16* aux_in;
    bf16* aux_out;
    const float* scale;
    void* D;
    int ldd, ldr, ld_aux;
    int act, dact, out_fp32;
    float alpha;
};

template <int BLOCK_N, int kStages>
struct SmemLayout {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
    static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024B alignment
};

// Apply the epilogue to 32 consecutive accumulator columns of one row and store them.
__device__ __forceinline__ void epilogue_store32(const EpiParams& ep, const uint32_t (&acc)[32], int row, int col0,
                                                 int ncols_valid) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * ep.alpha;

    if (ep.bias != nullptr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 8 < ncols_valid) {
                float b[8];
                unpack8(ldg16(ep.bias + col0 + c * 8), b);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[c * 8 + i] += b[i];
            }
        }
    }
    if (ep.aux_out != nullptr) {
        bf16* p = ep.aux_out + static_cast<size_t>(row) * ep.ld_aux + col0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 8 < ncols_valid) {
                float t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = v[c * 8 + i];
                stg16(p + c * 8, pack8(t));
            }
        }
    }
    if (ep.dact) {
        const bf16* p = ep.aux_in + static_cast<size_t>(row) * ep.ld_aux + col0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 8 < ncols_valid) {
                float u[8];
                unpack8(ldg16(p + c * 8), u);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[c * 8 + i] *= act_grad(ep.act, u[i]);
            }
        }
    } else if (ep.act != VLK_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = act_apply(ep.act, v[i]);
    }
    if (ep.scale != nullptr) {
        const float s = __ldg(ep.scale);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= s;
    }
    if (ep.residual != nullptr) {
        const bf16* p = ep.residual + static_cast<size_t>(row) * ep.ldr + col0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 8 < ncols_valid) {
                float r[8];
                unpack8(ldg16(p + c * 8), r);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[c * 8 + i] += r[i];
            }
        }
    }
    if (ep.out_fp32) {
        float* p = reinterpret_cast<float*>(ep.D) + static_cast<size_t>(row) * ep.ldd + col0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c * 4 < ncols_valid)
                *reinterpret_cast<float4*>(p + c * 4) = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
        }
    } else {
        bf16* p = reinterpret_cast<bf16*>(ep.D) + static_cast<size_t>(row) * ep.ldd + col0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 8 < ncols_valid) {
                float t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = v[c * 8 + i];
                stg16(p + c * 8, pack8(t));
            }
        }
    }
}

// A_MN / B_MN: operand is MN-major in global memory (see file header).
template <int BLOCK_N, int kStages, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         int M, int N, int K, EpiParams ep) {
    using L = SmemLayout<BLOCK_N, kStages>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries (descriptor base_offset = 0).
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    const int num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m_blocks * num_n_blocks;
    const int num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages; power of two >= 32

    if (warp_idx == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], kNumEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m_blocks;
                const int n_blk = tile / num_m_blocks;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes;
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                    if constexpr (!A_MN) {
                        // box = 64 (k) x 128 (m)
                        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    } else {
                        // two boxes of 64 (m) x 64 (k): each is one 8 KB swizzled slab
                        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], m_blk * BLOCK_M, kb * BLOCK_K);
                        ptx::tma_load_2d(sa + 8192, &tmap_a, &full_bar[stage], m_blk * BLOCK_M + 64, kb * BLOCK_K);
                    }
                    if constexpr (!B_MN) {
                        ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j)
                            ptx::tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n_blk * BLOCK_N + j * 64,
                                             kb * BLOCK_K);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =======================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int local_tile = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
                const int acc = local_tile & 1;
                const uint32_t acc_phase = (local_tile >> 1) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                ptx::tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after_sync();
                    const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t sb = sa + L::kABytes;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // K-major: 16 k-elements = 32 bytes inside the 128B swizzle row, rows grouped by 8
                        // (SBO = 1024 B).  MN-major: 16 k-rows = two 8-row groups of 1024 B (SBO), the next
                        // 64-wide MN slab is 8192 B away (LBO).
                        const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
                        ptx::umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
            }
        }
    } else if (warp_idx >= kEpiWarp0) {
        // ===================================== epilogue =========================================
        const int quad = warp_idx & 3;                        // TMEM lane quadrant this warp may access
        const int half = (warp_idx - kEpiWarp0) >> 2;         // which half of the tile's columns
        constexpr int kColsPerWarp = BLOCK_N / 2;
        int local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int m_blk = tile % num_m_blocks;
            const int n_blk = tile / num_m_blocks;
            const int acc = local_tile & 1;
            const uint32_t acc_phase = (local_tile >> 1) & 1;
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tc_fence_after_sync();
            const int row = m_blk * BLOCK_M + quad * 32 + lane;
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerWarp;
#pragma unroll 1
            for (int c = 0; c < kColsPerWarp; c += 32) {
                const int col0 = n_blk * BLOCK_N + half * kColsPerWarp + c;
                if (col0 >= N) break;  // warp-uniform
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(taddr + c, r);
                ptx::tmem_ld_wait();
                if (row < M) epilogue_store32(ep, r, row, col0, min(32, N - col0));
            }
            // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp_idx == 2) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 2D bf16 tensor map: `inner` contiguous elements per row, `outer` rows of stride ld elements;
// box = box_inner x box_outer, 128B swizzle, out-of-bounds reads return zero.
int make_tmap(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
              uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    VLK_REQUIRE(fn != nullptr, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLK_REQUIRE(r == CUDA_SUCCESS, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VLK_OK;
}

template <int BLOCK_N, int kStages, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, int sms,
           cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N, kStages>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, kStages, A_MN, B_MN>;
    static bool configured = false;  // per template instantiation
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
        configured = true;
    }
    const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
    const int grid = tiles < sms ? tiles : sms;
    kern<<<grid, kNumThreads, L::kDynamic, stream>>>(ta, tb, M, N, K, ep);
    VLK_CHECK_LAUNCH("vlk_gemm_bf16");
    return VLK_OK;
}

// Relative cost of finishing the problem with a given BLOCK_N: waves x per-tile time.  A 128xBN tile needs
// max(BN/2 tensor cycles, (128+BN)/4 smem-read cycles) per 16-deep k-step (B300_MICROARCH: floor = M*N/256,
// smem crossbar 128 B/clk).
double tile_cost(int M, int N, int bn, int sms) {
    const long tiles = static_cast<long>((M + BLOCK_M - 1) / BLOCK_M) * ((N + bn - 1) / bn);
    const long waves = (tiles + sms - 1) / sms;
    const double per_tile = fmax(bn / 2.0, (128.0 + bn) / 4.0) + 6.0;
    return waves * per_tile;
}

template <bool A_MN, bool B_MN>
int dispatch(const CUtensorMap* ta, const CUtensorMap* tb_by_bn /*[3]: 256,128,64*/, int M, int N, int K,
             const EpiParams& ep, int sms, int bn, cudaStream_t stream) {
    switch (bn) {
        case 256:
            return launch<256, 4, A_MN, B_MN>(*ta, tb_by_bn[0], M, N, K, ep, sms, stream);
        case 128:
            return launch<128, 6, A_MN, B_MN>(*ta, tb_by_bn[1], M, N, K, ep, sms, stream);
        default:
            return launch<64, 8, A_MN, B_MN>(*ta, tb_by_bn[2], M, N, K, ep, sms, stream);
    }
}

}  // namespace
}  // namespace vlk

using namespace vlk;

extern "C" int vlk_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                             int transA, int transB, const void* bias, const void* residual, int ldr,
                             const void* aux_in, void* aux_out, int ld_aux, const float* scale, int act, int dact,
                             float alpha, int out_fp32, void* stream) {
    VLK_REQUIRE(A && B && D, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: null operand");
    VLK_REQUIRE(M > 0 && N > 0 && K > 0, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
    VLK_REQUIRE(N % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: N=%d must be a multiple of 8", N);
    // K itself is free: K-major operands only need 16-byte row strides (lda/ldb % 8), MN-major operands carry K
    // as the outer TMA dimension; the K tail of the last 64-wide block is zero-filled by TMA.
    VLK_REQUIRE(transA || K <= lda, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: lda=%d < K=%d", lda, K);
    VLK_REQUIRE(transB || K <= ldb, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: ldb=%d < K=%d", ldb, K);
    VLK_REQUIRE(!transA || M % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: transA needs M %% 8 == 0 (M=%d)", M);
    VLK_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldd % 8 == 0, VLK_ERR_ALIGNMENT,
                "vlk_gemm_bf16: leading dims must be multiples of 8 (lda=%d ldb=%d ldd=%d)", lda, ldb, ldd);
    VLK_REQUIRE(aligned16(A) && aligned16(B) && aligned16(D), VLK_ERR_ALIGNMENT, "vlk_gemm_bf16: 16B alignment");
    VLK_REQUIRE(act >= VLK_ACT_NONE && act <= VLK_ACT_QUICK_GELU, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: act=%d", act);
    VLK_REQUIRE(!dact || aux_in, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16: dact needs aux_in");
    VLK_REQUIRE(!residual || (ldr % 8 == 0 && aligned16(residual)), VLK_ERR_ALIGNMENT, "vlk_gemm_bf16: residual");
    VLK_REQUIRE(!(aux_in || aux_out) || ld_aux % 8 == 0, VLK_ERR_ALIGNMENT, "vlk_gemm_bf16: ld_aux");
    VLK_REQUIRE(!bias || aligned16(bias), VLK_ERR_ALIGNMENT, "vlk_gemm_bf16: bias");
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_gemm_bf16: no sm_100 device");

    EpiParams ep;
    ep.bias = static_cast<const bf16*>(bias);
    ep.residual = static_cast<const bf16*>(residual);
    ep.aux_in = static_cast<const bf16*>(aux_in);
    ep.aux_out = static_cast<bf16*>(aux_out);
    ep.scale = scale;
    ep.D = D;
    ep.ldd = ldd;
    ep.ldr = ldr;
    ep.ld_aux = ld_aux;
    ep.act = act;
    ep.dact = dact;
    ep.out_fp32 = out_fp32;
    ep.alpha = alpha;

    // tile width: cheapest of 256 / 128 / 64 under the wave-quantisation model
    int bn = 256;
    double best = tile_cost(M, N, 256, sms);
    for (int cand : {128, 64}) {
        double c = tile_cost(M, N, cand, sms);
        if (c < best * 0.999) {
            best = c;
            bn = cand;
        }
    }
    if (const char* f = getenv("VLK_GEMM_BN")) {
        int v = atoi(f);
        if (v == 256 || v == 128 || v == 64) bn = v;
    }

    CUtensorMap ta, tb[3];
    int rc;
    if (!transA)
        rc = make_tmap(&ta, A, K, M, lda, BLOCK_K, BLOCK_M);
    else
        rc = make_tmap(&ta, A, M, K, lda, 64, BLOCK_K);
    if (rc) return rc;
    const int idx = bn == 256 ? 0 : (bn == 128 ? 1 : 2);
    if (!transB)
        rc = make_tmap(&tb[idx], B, K, N, ldb, BLOCK_K, bn);
    else
        rc = make_tmap(&tb[idx], B, N, K, ldb, 64, BLOCK_K);
    if (rc) return rc;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!transA && !transB) return dispatch<false, false>(&ta, tb, M, N, K, ep, sms, bn, s);
    if (!transA && transB) return dispatch<false, true>(&ta, tb, M, N, K, ep, sms, bn, s);
    if (transA && !transB) return dispatch<true, false>(&ta, tb, M, N, K, ep, sms, bn, s);
    return dispatch<true, true>(&ta, tb, M, N, K, ep, sms, bn, s);
}
