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
#include "gemm_internal.cuh"
#include "ptx.cuh"

namespace vlk {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kColVecBytes = 128 * 4 + 128 * 2;   // per epilogue warp: the folded-LayerNorm column sums and the bias of its columns

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
    int split_k;  // > 1: the K range is cut into split_k slices; each slice's partial tile goes to fp32 D either
                  // atomically (split_stride == 0, D zeroed by the caller) or into its own slab s * split_stride
    long long split_stride;
    // LayerNorm folded into the weights (vlk_gemm_bf16_lnfold): v = rstd[m] * (acc - mean[m] * colsum[n]), then bias
    const float* ln_mean;
    const float* ln_rstd;
    const float* ln_colsum;
    // ... or from the row sums (sum x, sum x^2) that the PRODUCING GEMM accumulated (stats_out of the residual GEMM
    // that wrote x): mean = s1 / K, rstd = rsqrt(s2 / K - mean^2 + eps)
    const float* ln_sums;
    float ln_inv_k, ln_eps;
    float* stats_out;   // [M][2] fp32, accumulated with atomicAdd: row sums of the bf16-ROUNDED outputs and of their squares
    // Fused lm_head + softmax cross-entropy (vlk_lmhead_ce_fwd / _bwd, lmhead_ce.cu).  The accumulator tile holds LOGITS:
    //   ce_mode 1 (forward): nothing is stored to D; every epilogue warp reduces its 32 rows x (BLOCK_N/2) logits to a
    //     running (max, sum exp) pair per row -> ce_partial[slice][row], slice = column / (BLOCK_N/2), and the thread
    //     that meets the row's label column writes that logit to ce_label_logit[row];
    //   ce_mode 2 (backward): D[m,n] = bf16((exp(logit - ce_lse[m]) - [n + ce_col0 == label[m]]) * ce_row_scale[m]),
    //     i.e. d loss / d logits of this vocabulary chunk, recomputed instead of read back.
    int ce_mode;
    const long long* ce_labels;
    float2* ce_partial;
    float* ce_label_logit;
    const float* ce_lse;
    const float* ce_row_scale;
    int ce_col0;
    int wide_io;  // every bf16 [M,N] operand of the epilogue (D, residual, aux) has 32-byte aligned rows: 256-bit accesses
#ifdef VLK_BRINGUP
    int debug;  // bring-up builds only (env VLK_GEMM_DEBUG): 1 = skip epilogue work, 2 = skip TMA loads and full-barrier waits
    unsigned long long* dbg_out;  // bring-up builds only: per-CTA wait-time breakdown (8 counters), see VLK_DBG_CLOCK
#endif
};

// Work-skipping switches exist in bring-up builds (-DVLK_BRINGUP) only: in the shipped library these fold to `false`.
#ifdef VLK_BRINGUP
#define VLK_DBG_SKIP_EPILOGUE(ep) (((ep).debug & 1) != 0)
#define VLK_DBG_SKIP_LOADS(ep) (((ep).debug & 2) != 0)
#define VLK_DBG_LDTM_ONLY(ep) (((ep).debug & 4) != 0)      // epilogue reads the accumulator and drops it
#define VLK_DBG_NO_GLOBAL_IO(ep) (((ep).debug & 8) != 0)   // epilogue computes but never stores
// cycle counters of the role threads: T(acc) { statement; } adds the cycles the statement took to acc
#define VLK_DBG_CLOCK_DECL(name) long long name = 0
#define VLK_DBG_TIMED(acc, stmt) do { const long long t0__ = clock64(); stmt; acc += clock64() - t0__; } while (0)
#define VLK_DBG_PUT(ep, slot, val) do { if ((ep).dbg_out) (ep).dbg_out[blockIdx.x * 8 + (slot)] = static_cast<unsigned long long>(val); } while (0)
unsigned long long* g_gemm_dbg_out = nullptr;
#else
#define VLK_DBG_SKIP_EPILOGUE(ep) false
#define VLK_DBG_SKIP_LOADS(ep) false
#define VLK_DBG_LDTM_ONLY(ep) false
#define VLK_DBG_NO_GLOBAL_IO(ep) false
#define VLK_DBG_CLOCK_DECL(name)
#define VLK_DBG_TIMED(acc, stmt) do { stmt; } while (0)
#define VLK_DBG_PUT(ep, slot, val) do { } while (0)
#endif

template <int BLOCK_N, int kStages>
struct SmemLayout {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kColVecOffset = kStages * kStageBytes;   // kNumEpiWarps x (128 fp32 colsum + 128 bf16 bias)
    static constexpr int kBarOffset = kColVecOffset + kNumEpiWarps * kColVecBytes;
    // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
    static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024B alignment
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ce_mode 2 on 32 logits of one row starting at chunk-local column col0: v <- (softmax - onehot) * row scale.
__device__ __forceinline__ void ce_grad32(const EpiParams& ep, float (&v)[32], int row, int col0) {
    constexpr float kLog2e = 1.4426950408889634f;
    const float nlse = -__ldg(ep.ce_lse + row) * kLog2e;
    const float w = __ldg(ep.ce_row_scale + row);
    const long long hit = ep.ce_labels[row] - ep.ce_col0 - col0;     // index of the label inside these 32 columns
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float p = ex2_approx(fmaf(v[i], kLog2e, nlse));
        v[i] = (p - (hit == i ? 1.0f : 0.0f)) * w;
    }
}

// ce_mode 1: one warp reduces its 32 rows x ncols_warp logits (TMEM) to (max, sum exp) per row and picks the label logit.
__device__ __forceinline__ void ce_stats_warp(const EpiParams& ep, uint32_t taddr, int row0, int n0, int ncols_warp, int M,
                                              int N, int lane) {
    constexpr float kLog2e = 1.4426950408889634f;
    const int row = row0 + lane;
    const long long label = row < M ? ep.ce_labels[row] : -1;
    float m = -INFINITY, s = 0.f;     // running max (in units of log2: logit * log2e) and sum of 2^(x - m)
#pragma unroll 1
    for (int c = 0; c < ncols_warp; c += 32) {
        const int col0 = n0 + c;
        if (col0 >= N) break;  // warp-uniform
        const int ncols = min(32, N - col0);
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + c, r);
        ptx::tmem_ld_wait();
        float v[32];
        float cm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v[i] = __uint_as_float(r[i]) * ep.alpha;
            if (i < ncols) cm = fmaxf(cm, v[i]);
        }
        const long long hit = label - col0;
        if (hit >= 0 && hit < ncols) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (hit == i) ep.ce_label_logit[row] = v[i];
        }
        const float nm = fmaxf(m, cm * kLog2e);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols) acc += ex2_approx(fmaf(v[i], kLog2e, -nm));
        s = s * ex2_approx(m - nm) + acc;     // first chunk: m = -inf -> factor 0
        m = nm;
    }
    if (row < M && n0 < N)
        ep.ce_partial[static_cast<size_t>(n0 / ncols_warp) * M + row] = make_float2(m, s);
}

// Apply the epilogue to 32 consecutive accumulator columns of one row and store them.
__device__ __forceinline__ void epilogue_store32(const EpiParams& ep, const uint32_t (&acc)[32], int row, int col0,
                                                 int ncols_valid, size_t d_off) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * ep.alpha;

    if (ep.ln_colsum != nullptr) {  // LayerNorm folded into the weights (narrow-tile path)
        float mu, rs;
        if (ep.ln_sums != nullptr) {
            const float2 s12 = __ldg(reinterpret_cast<const float2*>(ep.ln_sums) + row);
            mu = s12.x * ep.ln_inv_k;
            rs = rsqrtf(fmaxf(s12.y * ep.ln_inv_k - mu * mu, 0.f) + ep.ln_eps);
        } else {
            mu = __ldg(ep.ln_mean + row);
            rs = __ldg(ep.ln_rstd + row);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols_valid) v[i] = rs * (v[i] - mu * __ldg(ep.ln_colsum + col0 + i));
    }
    if (ep.ce_mode == 2) ce_grad32(ep, v, row, col0);
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
        float* p = reinterpret_cast<float*>(ep.D) + d_off + static_cast<size_t>(row) * ep.ldd + col0;
        if (ep.split_k > 1 && ep.split_stride == 0) {  // atomic split-K partial: D was zeroed by the caller
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (c < ncols_valid) atomicAdd(p + c, v[c]);
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c * 4 < ncols_valid)
                    *reinterpret_cast<float4*>(p + c * 4) =
                        make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
            }
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

// Which [M,N] input (if any) is read one chunk ahead of its use: the residual, else the activation-gradient input.
__device__ __forceinline__ const bf16* epi_staged_input(const EpiParams& ep, int& ld) {
    if (ep.residual != nullptr) {
        ld = ep.ldr;
        return ep.residual;
    }
    ld = ep.ld_aux;
    return ep.dact ? ep.aux_in : nullptr;
}

// ------------------------------------------------------------------------------------------------
// Direct epilogue: registers <-> global memory, no shared-memory staging.  After tcgen05.ld a thread holds 32
// consecutive columns of ONE row = 64 bytes of bf16 = two full 32-byte sectors, written with two 256-bit stores
// (STG.E.ENL2.256).  The [M,N] inputs (residual, activation-gradient input) are read the same way, one 32-column
// chunk ahead of its use (the first chunk before the accumulator wait).  Compared with the staged variant above this
// takes the epilogue's 2 x 64 KB (4 x with a staged input) per tile out of the shared-memory pipe that the tensor
// core's operand reads and the TMA writes already load to the limit, frees 64 KB for two more pipeline stages, and
// produces the pre-activation copy (aux_out) in the same pass.
// ------------------------------------------------------------------------------------------------
// Packed (f32x2: FFMA2 / FMUL2 / FADD2) forms of the forward activations for a pair of values: the epilogue warps are
// latency-bound at two warps per scheduler (IPC 0.15-0.27 per scheduler under ncu), so what shortens a tile's epilogue
// is fewer instructions per element, not a faster pipe.
__device__ __forceinline__ void act_apply2(int act, float& x0, float& x1) {
    if (act == VLK_ACT_GELU_TANH) {   // 0.5 x (1 + tanh(k0 (x + k1 x^3))) = x (0.5 tanh(x (k0 + k0 k1 x^2)) + 0.5)
        constexpr float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
        float s0, s1, w0, w1;
        ptx::fmul2(s0, s1, x0, x1, x0, x1);
        ptx::ffma2(w0, w1, s0, s1, k01, k01, k0, k0);
        ptx::fmul2(w0, w1, w0, w1, x0, x1);
        w0 = tanh_fast(w0);
        w1 = tanh_fast(w1);
        ptx::ffma2(w0, w1, w0, w1, 0.5f, 0.5f, 0.5f, 0.5f);
        ptx::fmul2(x0, x1, x0, x1, w0, w1);
    } else if (act == VLK_ACT_QUICK_GELU) {   // x sigmoid(1.702 x) = x (0.5 tanh(0.851 x) + 0.5)
        float w0, w1;
        ptx::fmul2(w0, w1, x0, x1, 0.851f, 0.851f);
        w0 = tanh_fast(w0);
        w1 = tanh_fast(w1);
        ptx::ffma2(w0, w1, w0, w1, 0.5f, 0.5f, 0.5f, 0.5f);
        ptx::fmul2(x0, x1, x0, x1, w0, w1);
    } else if (act != VLK_ACT_NONE) {
        x0 = act_apply(act, x0);
        x1 = act_apply(act, x1);
    }
}

struct EpiPre {
    uint32_t r[16];  // 32 bf16 of this thread's row
};

__device__ __forceinline__ void ldg_row32(uint32_t (&r)[16], const bf16* p, int ncols, bool wide) {
    if (wide && ncols == 32) {
        ptx::ldg_v8(r, p);
        ptx::ldg_v8(r + 8, p + 16);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (q * 8 < ncols) u = ldg16(p + q * 8);
            r[q * 4] = u.x, r[q * 4 + 1] = u.y, r[q * 4 + 2] = u.z, r[q * 4 + 3] = u.w;
        }
    }
}
__device__ __forceinline__ void stg_row32(bf16* p, const uint32_t (&r)[16], int ncols, bool wide) {
    if (wide && ncols == 32) {
        ptx::stg_v8(p, r);
        ptx::stg_v8(p + 16, r + 8);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q * 8 < ncols) stg16(p + q * 8, make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]));
    }
}
__device__ __forceinline__ void pack32(const float (&v)[32], uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bf162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        r[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const bf162*>(&u)); }

// The column vectors of the warp's tile columns (bias; column sums of a folded LayerNorm) are staged in shared memory
// BEFORE the accumulator wait, one coalesced load per warp and tile.  Read with __ldg inside the chunk loop they were
// the epilogue's critical path: every 32-column chunk waited a full L2 round trip for them (ncu: 43 % of the kernel's
// stall samples were long-scoreboard waits on their first use) and the MMA warp waited 30-45 % of its time for a free
// accumulator on every K <= 1024 product.
struct ColVec {
    const float* colsum;   // [ncols_warp] fp32 (valid when the launch folds a LayerNorm)
    const bf16* bias;      // [ncols_warp] bf16 (valid when the launch has a bias)
    float ln_mu, ln_rs;    // this thread's row statistics of a folded LayerNorm (fetched under the MMAs as well)
};
__device__ __forceinline__ ColVec stage_colvec(const EpiParams& ep, uint8_t* buf, int row, int M, int n0, int ncols_warp,
                                               int N, int lane) {
    float* cs = reinterpret_cast<float*>(buf);
    bf16* bs = reinterpret_cast<bf16*>(buf + 128 * 4);
    float ln_mu = 0.f, ln_rs = 1.f;
    if (ep.ln_colsum != nullptr && row < M) {
        if (ep.ln_sums != nullptr) {
            const float2 s12 = __ldg(reinterpret_cast<const float2*>(ep.ln_sums) + row);
            ln_mu = s12.x * ep.ln_inv_k;
            ln_rs = rsqrtf(fmaxf(s12.y * ep.ln_inv_k - ln_mu * ln_mu, 0.f) + ep.ln_eps);
        } else {
            ln_mu = __ldg(ep.ln_mean + row);
            ln_rs = __ldg(ep.ln_rstd + row);
        }
    }
    __syncwarp();   // the previous tile's readers are done
    for (int c = lane * 4; c < ncols_warp; c += 128) {
        const bool ok = n0 + c < N;   // N % 8 == 0 and c % 4 == 0: a group of four never straddles N
        if (ep.ln_colsum != nullptr)
            *reinterpret_cast<float4*>(cs + c) =
                ok ? __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + n0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias != nullptr)
            *reinterpret_cast<uint2*>(bs + c) =
                ok ? __ldg(reinterpret_cast<const uint2*>(ep.bias + n0 + c)) : make_uint2(0u, 0u);
    }
    __syncwarp();
    return ColVec{cs, bs, ln_mu, ln_rs};
}

// The first chunk of the staged input, fetched BEFORE the accumulator wait (the warp is idle while its tile's MMAs run).
__device__ __forceinline__ void epilogue_prefetch_direct(const EpiParams& ep, EpiPre& pre, int row0, int n0, int M, int N,
                                                         int lane) {
    if (ep.out_fp32 || ep.ce_mode == 1 || VLK_DBG_SKIP_EPILOGUE(ep) || n0 >= N) return;
    int ld;
    const bf16* in = epi_staged_input(ep, ld);
    const int row = row0 + lane;
    if (in == nullptr || row >= M) return;
    ldg_row32(pre.r, in + static_cast<size_t>(row) * ld + n0, min(32, N - n0), ep.wide_io != 0);
}

template <int ACT, int DACT, int RES, int SCALE, int AUX, int LNF, int STATS = 0, int CE = 0>
__device__ __forceinline__ void epilogue_direct_t(const EpiParams& ep, EpiPre& pre, const ColVec& cv, uint32_t taddr,
                                                  int row0, int n0, int ncols_warp, int M, int N, int lane) {
    const int row = row0 + lane;
    const bool row_ok = row < M;
    const int act = ACT < 0 ? ep.act : ACT;
    const bool dact = DACT < 0 ? (ep.dact != 0) : (DACT != 0);
    const bool has_res = RES < 0 ? (ep.residual != nullptr) : (RES != 0);
    const bool has_scale = SCALE < 0 ? (ep.scale != nullptr) : (SCALE != 0);
    const bool has_aux_out = AUX < 0 ? (ep.aux_out != nullptr) : (AUX != 0);
    const bool ln_fold = LNF < 0 ? (ep.ln_colsum != nullptr) : (LNF != 0);
    const bool stats = STATS < 0 ? (ep.stats_out != nullptr) : (STATS != 0);
    const bool wide = ep.wide_io != 0;
    const float ln_mu = cv.ln_mu, ln_rs = cv.ln_rs;
    float st1 = 0.f, st2 = 0.f;
    const bool aux_direct = dact && has_res;  // both present: the residual is the prefetched stream, aux_in is read in place
    const float scale = has_scale ? __ldg(ep.scale) : 1.0f;
    int ld_in = 0;
    const bf16* in = (has_res || dact) ? epi_staged_input(ep, ld_in) : nullptr;
    const size_t in_row = static_cast<size_t>(row) * ld_in;
#pragma unroll 1
    for (int c = 0; c < ncols_warp; c += 32) {
        const int col0 = n0 + c;
        if (col0 >= N) break;  // warp-uniform
        const int ncols = min(32, N - col0);
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + c, r);
        EpiPre nxt;
        const bool has_next = in != nullptr && c + 32 < ncols_warp && col0 + 32 < N;
        if (has_next && row_ok) ldg_row32(nxt.r, in + in_row + col0 + 32, min(32, N - col0 - 32), wide);
        ptx::tmem_ld_wait();
        if (VLK_DBG_LDTM_ONLY(ep)) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (ep.alpha != 1.0f) {   // uniform
#pragma unroll
            for (int i = 0; i < 32; i += 2) ptx::fmul2(v[i], v[i + 1], v[i], v[i + 1], ep.alpha, ep.alpha);
        }
        if (CE != 0) {
            if (row_ok) ce_grad32(ep, v, row, col0);
        }
        if (ln_fold) {   // v = rstd (v - mean colsum) = v rstd + colsum (-rstd mean)
            const float nrm = -ln_rs * ln_mu;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q * 4 < ncols) {
                    const float4 c4 = *reinterpret_cast<const float4*>(cv.colsum + c + q * 4);   // broadcast read
                    float t0, t1, t2, t3;
                    ptx::fmul2(t0, t1, c4.x, c4.y, nrm, nrm);
                    ptx::fmul2(t2, t3, c4.z, c4.w, nrm, nrm);
                    ptx::ffma2(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 0], v[q * 4 + 1], ln_rs, ln_rs, t0, t1);
                    ptx::ffma2(v[q * 4 + 2], v[q * 4 + 3], v[q * 4 + 2], v[q * 4 + 3], ln_rs, ln_rs, t2, t3);
                }
            }
        }
        if (ep.bias != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q * 8 < ncols) {
                    const uint4 bq = *reinterpret_cast<const uint4*>(cv.bias + c + q * 8);   // broadcast read
                    const float2 b0 = unpack2(bq.x), b1 = unpack2(bq.y), b2 = unpack2(bq.z), b3 = unpack2(bq.w);
                    ptx::fadd2(v[q * 8 + 0], v[q * 8 + 1], v[q * 8 + 0], v[q * 8 + 1], b0.x, b0.y);
                    ptx::fadd2(v[q * 8 + 2], v[q * 8 + 3], v[q * 8 + 2], v[q * 8 + 3], b1.x, b1.y);
                    ptx::fadd2(v[q * 8 + 4], v[q * 8 + 5], v[q * 8 + 4], v[q * 8 + 5], b2.x, b2.y);
                    ptx::fadd2(v[q * 8 + 6], v[q * 8 + 7], v[q * 8 + 6], v[q * 8 + 7], b3.x, b3.y);
                }
            }
        }
        uint32_t pk[16];
        if (has_aux_out) {  // the pre-activation copy the backward pass needs, from the same registers
            pack32(v, pk);
            if (row_ok) stg_row32(ep.aux_out + static_cast<size_t>(row) * ep.ld_aux + col0, pk, ncols, wide);
        }
        if (dact) {
            if (aux_direct) {
                uint32_t u[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = 0;
                if (row_ok) ldg_row32(u, ep.aux_in + static_cast<size_t>(row) * ep.ld_aux + col0, ncols, wide);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 t = unpack2(u[i]);
                    v[2 * i] *= act_grad(act, t.x);
                    v[2 * i + 1] *= act_grad(act, t.y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 t = unpack2(pre.r[i]);
                    v[2 * i] *= act_grad(act, t.x);
                    v[2 * i + 1] *= act_grad(act, t.y);
                }
            }
        } else if (act != VLK_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) act_apply2(act, v[i], v[i + 1]);
        }
        if (has_scale) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) ptx::fmul2(v[i], v[i + 1], v[i], v[i + 1], scale, scale);
        }
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 t = unpack2(pre.r[i]);
                ptx::fadd2(v[2 * i], v[2 * i + 1], v[2 * i], v[2 * i + 1], t.x, t.y);
            }
        }
        pack32(v, pk);
        if (stats) {  // statistics of what the consumer will actually read (the rounded values)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (2 * i < ncols) {
                    const float2 t = unpack2(pk[i]);
                    st1 += t.x + t.y;
                    st2 = fmaf(t.x, t.x, fmaf(t.y, t.y, st2));
                }
            }
        }
        if (row_ok && !(VLK_DBG_NO_GLOBAL_IO(ep) && pk[3] != 0x12345678u))
            stg_row32(reinterpret_cast<bf16*>(ep.D) + static_cast<size_t>(row) * ep.ldd + col0, pk, ncols, wide);
        if (has_next) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pre.r[i] = nxt.r[i];
        }
    }
    if (stats && row_ok) {
        atomicAdd(ep.stats_out + 2 * static_cast<size_t>(row), st1);
        atomicAdd(ep.stats_out + 2 * static_cast<size_t>(row) + 1, st2);
    }
}

template <bool SPECIALISE>
__device__ __forceinline__ void epilogue_warp_direct(const EpiParams& ep, EpiPre& pre, const ColVec& cv, uint32_t taddr,
                                                     int row0, int n0, int ncols_warp, int M, int N, int lane,
                                                     size_t d_off = 0) {
    if (VLK_DBG_SKIP_EPILOGUE(ep)) return;
    if (ep.ce_mode == 1) {
        ce_stats_warp(ep, taddr, row0, n0, ncols_warp, M, N, lane);
        return;
    }
    if (ep.out_fp32) {  // fp32 output (split-K slabs): row-per-thread float4 stores
        const int row = row0 + lane;
#pragma unroll 1
        for (int c = 0; c < ncols_warp; c += 32) {
            const int col0 = n0 + c;
            if (col0 >= N) break;
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(taddr + c, r);
            ptx::tmem_ld_wait();
            if (row < M) epilogue_store32(ep, r, row, col0, min(32, N - col0), d_off);
        }
        return;
    }
#define VLK_EPI(ACT, DACT, RES, SCALE, AUX) \
    epilogue_direct_t<ACT, DACT, RES, SCALE, AUX, 0>(ep, pre, cv, taddr, row0, n0, ncols_warp, M, N, lane)
#define VLK_EPI_LN(ACT) epilogue_direct_t<ACT, 0, 0, 0, 0, 1>(ep, pre, cv, taddr, row0, n0, ncols_warp, M, N, lane)
    const bool res = ep.residual != nullptr, sc = ep.scale != nullptr, aux = ep.aux_out != nullptr;
    if (ep.ce_mode == 2) {   // d logits of a vocabulary chunk
        epilogue_direct_t<0, 0, 0, 0, 0, 0, 0, 1>(ep, pre, cv, taddr, row0, n0, ncols_warp, M, N, lane);
        return;
    }
    if constexpr (!SPECIALISE) {
        epilogue_direct_t<-1, -1, -1, -1, -1, -1, -1>(ep, pre, cv, taddr, row0, n0, ncols_warp, M, N, lane);
        return;
    }
    if (ep.stats_out != nullptr) {  // residual GEMM that also produces the row statistics of its output
        epilogue_direct_t<0, 0, -1, 0, 0, 0, 1>(ep, pre, cv, taddr, row0, n0, ncols_warp, M, N, lane);
    } else if (ep.ln_colsum != nullptr) {   // LayerNorm folded into the weights: plain / quick-GELU / tanh-GELU bodies
        if (ep.act == VLK_ACT_QUICK_GELU) VLK_EPI_LN(VLK_ACT_QUICK_GELU);
        else if (ep.act == VLK_ACT_GELU_TANH) VLK_EPI_LN(VLK_ACT_GELU_TANH);
        else VLK_EPI_LN(VLK_ACT_NONE);
    } else if (!sc && !ep.dact && ep.act == VLK_ACT_NONE && !aux) {          // (bias) [+ residual]: projections, dgrad, wgrad
        if (res) VLK_EPI(0, 0, 1, 0, 0);
        else VLK_EPI(0, 0, 0, 0, 0);
    } else if (!sc && !ep.dact && !res && ep.act == VLK_ACT_GELU_TANH) {   // GPT-2 c_fc (+ saved pre-activation)
        VLK_EPI(VLK_ACT_GELU_TANH, 0, 0, 0, -1);
    } else if (!sc && !ep.dact && !res && ep.act == VLK_ACT_QUICK_GELU) {  // CLIP fc1
        VLK_EPI(VLK_ACT_QUICK_GELU, 0, 0, 0, -1);
    } else if (!sc && ep.dact && !res && !aux && ep.act == VLK_ACT_GELU_TANH) {  // GPT-2 c_proj dgrad x gelu'
        VLK_EPI(VLK_ACT_GELU_TANH, 1, 0, 0, 0);
    } else {                                                          // everything else (erf-GELU, gated residual, ...)
        VLK_EPI(-1, -1, -1, -1, -1);
    }
#undef VLK_EPI
#undef VLK_EPI_LN
}


// Rasterisation of work units onto the output grid.  A wave of ~148 concurrently running CTAs should touch as few
// distinct A and B tiles as possible, because what bounds this kernel is unique bytes leaving the L2 (operands
// shared by concurrent CTAs are merged there).  Units are therefore walked in bands of `group` row-units: inside
// a band the row index is fastest, then the column block, so a wave covers a (group x 148/group) patch of tiles.
__device__ __forceinline__ void unit_to_mn(int u, int num_m_units, int num_n_blocks, int group, int& m_unit,
                                           int& n_blk) {
    const int band_units = group * num_n_blocks;
    const int band = u / band_units;
    const int rem = u - band * band_units;
    const int rows = min(group, num_m_units - band * group);  // the last band may be shorter
    n_blk = rem / rows;
    m_unit = band * group + (rem - n_blk * rows);
}

// A_MN / B_MN: operand is MN-major in global memory (see file header).
// CLUSTER: CTAs per thread-block cluster.  The CTAs of a cluster own CLUSTER consecutive 128-row blocks of the
// same column block, so they need the same B tile: each CTA fetches 1/CLUSTER of it and TMA-multicasts the slice
// into every CTA's shared memory.  L2->SM traffic per CTA and k-block drops from (128 + BN) to (128 + BN/CLUSTER)
// rows of 128 B, which is what bounds this kernel (one SM needs ~94 B/clk un-shared, the L2 delivers ~45).
template <int BLOCK_N, int kStages, bool A_MN, bool B_MN, int CLUSTER>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         int M, int N, int K, int group, EpiParams ep) {
    using L = SmemLayout<BLOCK_N, kStages>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries (descriptor base_offset = 0).
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;

    // Work unit = CLUSTER vertically adjacent tiles; cluster c walks units c, c + num_clusters, ...  The last unit
    // of a column may hang over the bottom edge: such a CTA runs the full protocol on zero-filled rows.
    const int num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    const int num_m_units = (num_m_blocks + CLUSTER - 1) / CLUSTER;
    const int num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_out_tiles = num_m_units * num_n_blocks;
    const int num_tiles = num_out_tiles * ep.split_k;  // split-K: slice index is the slowest-varying part
    const int total_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per_split = (total_k_blocks + ep.split_k - 1) / ep.split_k;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages; power of two >= 32
    const int cta_rank = CLUSTER > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
    const int unit0 = blockIdx.x / CLUSTER, unit_stride = gridDim.x / CLUSTER;
    constexpr uint16_t kMcastMask = static_cast<uint16_t>((1u << CLUSTER) - 1);

    if (warp_idx == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], CLUSTER);  // released by the MMA warp of every CTA in the cluster
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
    if (CLUSTER > 1) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        // The warp walks the loop converged; one elected lane issues (see the MMA issuer below).
        const bool issuer = ptx::elect_one();
        if (!VLK_DBG_SKIP_LOADS(ep)) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
                int m_unit, n_blk;
                unit_to_mn(tile % num_out_tiles, num_m_units, num_n_blocks, group, m_unit, n_blk);
                const int m_blk = m_unit * CLUSTER + cta_rank;
                const int kb0 = (tile / num_out_tiles) * kb_per_split, kb1 = min(kb0 + kb_per_split, total_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes;
                    if (issuer) {
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                    if constexpr (!A_MN) {
                        // box = 64 (k) x 128 (m)
                        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    } else {
                        // two boxes of 64 (m) x 64 (k): each is one 8 KB swizzled slab
                        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], m_blk * BLOCK_M, kb * BLOCK_K);
                        ptx::tma_load_2d(sa + 8192, &tmap_a, &full_bar[stage], m_blk * BLOCK_M + 64, kb * BLOCK_K);
                    }
                    if constexpr (CLUSTER == 1) {
                        if constexpr (!B_MN) {
                            ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
                        } else {
#pragma unroll
                            for (int j = 0; j < BLOCK_N / 64; ++j)
                                ptx::tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], n_blk * BLOCK_N + j * 64,
                                                 kb * BLOCK_K);
                        }
                    } else {
                        // this CTA's slice of the shared B tile, multicast to the whole cluster
                        constexpr int kSliceRows = BLOCK_N / CLUSTER;
                        const int n0 = n_blk * BLOCK_N + cta_rank * kSliceRows;
                        uint8_t* dst = sb + cta_rank * kSliceRows * 128;
                        if constexpr (!B_MN) {
                            ptx::tma_load_2d_mcast(dst, &tmap_b, &full_bar[stage], kb * BLOCK_K, n0, kMcastMask);
                        } else {
#pragma unroll
                            for (int j = 0; j < kSliceRows / 64; ++j)
                                ptx::tma_load_2d_mcast(dst + j * 8192, &tmap_b, &full_bar[stage], n0 + j * 64,
                                                       kb * BLOCK_K, kMcastMask);
                        }
                    }
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =======================================
        // The whole warp walks the loop CONVERGED (every lane polls the barriers) and one elected lane issues, so every
        // operand of tcgen05.mma is warp-uniform and lives in uniform registers; the shared-memory descriptors are a base
        // plus a stage / k-step offset in the address field.  Issued from inside `if (lane == 0)` the same loop compiled
        // to ~25 instructions per MMA (an ELECT / R2UR.BROADCAST loop per operand set): ~600 clk per 64-deep k-block
        // against the 512 clk the tensor core needs for it — the ISSUE loop, not the tensor pipe, bounded the kernel
        // (profiles/r02/gemm_issue_loop.md).
        {
            const bool issuer = ptx::elect_one();
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
            // K-major: 16 k-elements = 32 bytes inside the 128B swizzle row, rows grouped by 8 (SBO = 1024 B).
            // MN-major: 16 k-rows = two 8-row groups of 1024 B (SBO), the next 64-wide MN slab is 8192 B away (LBO).
            const uint32_t smem_base = ptx::smem_u32(smem);
            const uint64_t da0 = A_MN ? ptx::make_smem_desc_sw128(smem_base, 8192, 1024)
                                      : ptx::make_smem_desc_sw128(smem_base, 16, 1024);
            const uint64_t db0 = B_MN ? ptx::make_smem_desc_sw128(smem_base + L::kABytes, 8192, 1024)
                                      : ptx::make_smem_desc_sw128(smem_base + L::kABytes, 16, 1024);
            constexpr uint32_t kStepA = (A_MN ? 2048 : 32) >> 4, kStepB = (B_MN ? 2048 : 32) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int local_tile = 0;
            for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++local_tile) {
                const int acc = local_tile & 1;
                const uint32_t acc_phase = (local_tile >> 1) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                ptx::tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                const int kb0 = (tile / num_out_tiles) * kb_per_split, kb1 = min(kb0 + kb_per_split, total_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!VLK_DBG_SKIP_LOADS(ep)) ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after_sync();
                    const uint32_t soff = static_cast<uint32_t>(stage) * (L::kStageBytes >> 4);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            ptx::umma_bf16_ss(tmem_d, da0 + (soff + k * kStepA), db0 + (soff + k * kStepB), idesc,
                                              ((kb - kb0) | k) != 0 ? 1u : 0u);
                        // frees the smem slot (in every CTA that multicasts into it) when these MMAs retire
                        if constexpr (CLUSTER == 1) ptx::umma_commit(&empty_bar[stage]);
                        else ptx::umma_commit_mcast(&empty_bar[stage], kMcastMask);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (issuer) ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
                __syncwarp();
            }
        }
    } else if (warp_idx >= kEpiWarp0) {
        // ===================================== epilogue =========================================
        const int quad = warp_idx & 3;                        // TMEM lane quadrant this warp may access
        const int half = (warp_idx - kEpiWarp0) >> 2;         // which half of the tile's columns
        constexpr int kColsPerWarp = BLOCK_N / 2;
        int local_tile = 0;
        for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++local_tile) {
            int m_unit, n_blk;
            unit_to_mn(tile % num_out_tiles, num_m_units, num_n_blocks, group, m_unit, n_blk);
            const int m_blk = m_unit * CLUSTER + cta_rank;
            const int acc = local_tile & 1;
            const uint32_t acc_phase = (local_tile >> 1) & 1;
            const int row0 = m_blk * BLOCK_M + quad * 32, n0 = n_blk * BLOCK_N + half * kColsPerWarp;
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerWarp;
            EpiPre pre;
            epilogue_prefetch_direct(ep, pre, row0, n0, M, N, lane);
            const ColVec cv = stage_colvec(ep, smem + L::kColVecOffset + (warp_idx - kEpiWarp0) * kColVecBytes, row0 + lane, M,
                                           n0, kColsPerWarp, N, lane);
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tc_fence_after_sync();
            epilogue_warp_direct<false>(ep, pre, cv, taddr, row0, n0, kColsPerWarp, M, N, lane,
                                        static_cast<size_t>(ep.split_stride) * (tile / num_out_tiles));
            // all TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_relaxed(&tmem_empty_bar[acc]);
        }
    }

    ptx::tc_fence_before_sync();
    // no CTA may leave while a peer can still multicast into its shared memory or arrive on its barriers
    if (CLUSTER > 1) ptx::cluster_sync_all(); else __syncthreads();
    if (warp_idx == 2) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// cta_group::2 variant: a CTA pair (two SMs of one TPC) owns a 256 x BLOCK_N output tile.  Each CTA stages its own
// 128 rows of A and its own BLOCK_N/2 rows of B; ONE tcgen05.mma issued by the leader drives both tensor cores
// (M = 256), each SM reading the other half of B through the pair's operand-exchange path.  Per SM this halves the
// B bytes written by TMA and read by the tensor core: shared-memory traffic drops from 2 x 96 to 2 x 64 B/clk (the
// single-CTA 128x256 tile is bound by the 128 B/clk shared-memory port, not by the tensor pipe) and the smaller
// stage (32 KB instead of 48 KB) buys a 6-deep ring.
//
// Barrier topology (L = leader = even cluster rank, P = peer):
//   full[s]       lives in L; armed by L's producer with the bytes of BOTH CTAs; both CTAs' TMA credit it.
//   empty[s]      one in each CTA; released for both by L's tcgen05.commit (multicast).
//   tmem_full[a]  one in each CTA; signalled for both by L's commit; each CTA's epilogue drains its own TMEM half.
//   tmem_empty[a] lives in L; 2 x kNumEpiWarps arrivals (P's epilogue warps arrive remotely).
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int kStages>
struct SmemLayout2 {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = (BLOCK_N / 2) * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kColVecOffset = kStages * kStageBytes;
    static constexpr int kBarOffset = kColVecOffset + kNumEpiWarps * kColVecBytes;
    static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;
};

template <int BLOCK_N, int kStages, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M,
                      int N, int K, int group, EpiParams ep) {
    using L = SmemLayout2<BLOCK_N, kStages>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int cta_rank = static_cast<int>(ptx::cluster_ctarank());
    const bool is_leader = cta_rank == 0;

    const int num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    const int num_m_units = (num_m_blocks + 1) / 2;
    const int num_n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_out_tiles = num_m_units * num_n_blocks;
    const int num_tiles = num_out_tiles * ep.split_k;
    const int total_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    const int kb_per_split = (total_k_blocks + ep.split_k - 1) / ep.split_k;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    const int unit0 = blockIdx.x / 2, unit_stride = gridDim.x / 2;

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
            ptx::mbar_init(&tmem_empty_bar[s], 2 * kNumEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    if (warp_idx == 2) {
        ptx::tmem_alloc_2sm(tmem_slot, kTmemCols);
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before_sync();
    ptx::cluster_sync_all();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        // ===================================== TMA producer (both CTAs) =========================
        const bool issuer = ptx::elect_one();
        if (!VLK_DBG_SKIP_LOADS(ep)) {
            int stage = 0;
            uint32_t phase = 0;
            VLK_DBG_CLOCK_DECL(c_slot);
            for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
                int m_unit, n_blk;
                unit_to_mn(tile % num_out_tiles, num_m_units, num_n_blocks, group, m_unit, n_blk);
                const int m_blk = m_unit * 2 + cta_rank;
                const int n0 = n_blk * BLOCK_N + cta_rank * (BLOCK_N / 2);
                const int kb0 = (tile / num_out_tiles) * kb_per_split, kb1 = min(kb0 + kb_per_split, total_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    VLK_DBG_TIMED(c_slot, ptx::mbar_wait(&empty_bar[stage], phase ^ 1));
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes;
                    if (issuer) {
                    if (is_leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    if constexpr (!A_MN) {
                        ptx::tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
                    } else {
                        ptx::tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], m_blk * BLOCK_M, kb * BLOCK_K);
                        ptx::tma_load_2d_2sm(sa + 8192, &tmap_a, &full_bar[stage], m_blk * BLOCK_M + 64, kb * BLOCK_K);
                    }
                    if constexpr (!B_MN) {
                        ptx::tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 128; ++j)
                            ptx::tma_load_2d_2sm(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + j * 64, kb * BLOCK_K);
                    }
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (issuer) VLK_DBG_PUT(ep, 0, c_slot);
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer (leader CTA only) =====================
        // Converged warp, one elected lane issues, operands in uniform registers (see the single-CTA kernel).
        if (is_leader) {
            const bool issuer = ptx::elect_one();
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(2 * BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
            const uint32_t smem_base = ptx::smem_u32(smem);
            const uint64_t da0 = A_MN ? ptx::make_smem_desc_sw128(smem_base, 8192, 1024)
                                      : ptx::make_smem_desc_sw128(smem_base, 16, 1024);
            const uint64_t db0 = B_MN ? ptx::make_smem_desc_sw128(smem_base + L::kABytes, 8192, 1024)
                                      : ptx::make_smem_desc_sw128(smem_base + L::kABytes, 16, 1024);
            constexpr uint32_t kStepA = (A_MN ? 2048 : 32) >> 4, kStepB = (B_MN ? 2048 : 32) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int local_tile = 0;
            VLK_DBG_CLOCK_DECL(c_acc);
            VLK_DBG_CLOCK_DECL(c_full);
            VLK_DBG_CLOCK_DECL(c_total);
#ifdef VLK_BRINGUP
            c_total = -clock64();
#endif
            for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++local_tile) {
                const int acc = local_tile & 1;
                const uint32_t acc_phase = (local_tile >> 1) & 1;
                VLK_DBG_TIMED(c_acc, ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1));
                ptx::tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                const int kb0 = (tile / num_out_tiles) * kb_per_split, kb1 = min(kb0 + kb_per_split, total_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!VLK_DBG_SKIP_LOADS(ep)) VLK_DBG_TIMED(c_full, ptx::mbar_wait(&full_bar[stage], phase));
                    ptx::tc_fence_after_sync();
                    const uint32_t soff = static_cast<uint32_t>(stage) * (L::kStageBytes >> 4);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            ptx::umma_bf16_ss_2sm(tmem_d, da0 + (soff + k * kStepA), db0 + (soff + k * kStepB), idesc,
                                                  ((kb - kb0) | k) != 0 ? 1u : 0u);
                        ptx::umma_commit_2sm(&empty_bar[stage], 0b11);  // frees the slot in both CTAs
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (issuer) ptx::umma_commit_2sm(&tmem_full_bar[acc], 0b11);  // both halves of the accumulator are complete
                __syncwarp();
            }
#ifdef VLK_BRINGUP
            c_total += clock64();
#endif
            if (issuer) {
                VLK_DBG_PUT(ep, 1, c_acc);
                VLK_DBG_PUT(ep, 2, c_full);
                VLK_DBG_PUT(ep, 3, c_total);
            }
        }
    } else if (warp_idx >= kEpiWarp0) {
        // ===================================== epilogue (both CTAs, own 128 rows) ===============
        const int quad = warp_idx & 3;
        const int half = (warp_idx - kEpiWarp0) >> 2;
        constexpr int kColsPerWarp = BLOCK_N / 2;
        int local_tile = 0;
        VLK_DBG_CLOCK_DECL(c_pre);
        VLK_DBG_CLOCK_DECL(c_wait);
        VLK_DBG_CLOCK_DECL(c_work);
        for (int tile = unit0; tile < num_tiles; tile += unit_stride, ++local_tile) {
            int m_unit, n_blk;
            unit_to_mn(tile % num_out_tiles, num_m_units, num_n_blocks, group, m_unit, n_blk);
            const int m_blk = m_unit * 2 + cta_rank;
            const int acc = local_tile & 1;
            const uint32_t acc_phase = (local_tile >> 1) & 1;
            const int row0 = m_blk * BLOCK_M + quad * 32, n0 = n_blk * BLOCK_N + half * kColsPerWarp;
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerWarp;
            EpiPre pre;
            VLK_DBG_TIMED(c_pre, epilogue_prefetch_direct(ep, pre, row0, n0, M, N, lane));
            const ColVec cv = stage_colvec(ep, smem + L::kColVecOffset + (warp_idx - kEpiWarp0) * kColVecBytes, row0 + lane, M,
                                           n0, kColsPerWarp, N, lane);
            VLK_DBG_TIMED(c_wait, ptx::mbar_wait(&tmem_full_bar[acc], acc_phase));
            ptx::tc_fence_after_sync();
            VLK_DBG_TIMED(c_work, epilogue_warp_direct<(BLOCK_N == 256)>(ep, pre, cv, taddr, row0, n0, kColsPerWarp, M, N, lane,
                          static_cast<size_t>(ep.split_stride) * (tile / num_out_tiles)));
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(&tmem_empty_bar[acc], 0);  // the leader's barrier
        }
        if (warp_idx == kEpiWarp0 && lane == 0) {
            VLK_DBG_PUT(ep, 4, c_pre);
            VLK_DBG_PUT(ep, 5, c_wait);
            VLK_DBG_PUT(ep, 6, c_work);
        }
    }

    ptx::tc_fence_before_sync();
    ptx::cluster_sync_all();
    if (warp_idx == 2) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
}

// Split-K second pass: out[r][c] = (accumulate ? out[r][c] : 0) + sum_s ws[s][r][c], rounded to bf16 once.
// The slabs were just written by the GEMM and are L2-resident (a few MB); 8 columns per thread.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long slab, bf16* __restrict__ out, int M, int N,
                     int ldd, int accumulate) {
    const int n8 = N >> 3;
    const long long total = static_cast<long long>(M) * n8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / n8), c = static_cast<int>(i - static_cast<long long>(r) * n8) * 8;
        const float* src = ws + static_cast<size_t>(r) * N + c;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bf16* dst = out + static_cast<size_t>(r) * ldd + c;
        if (accumulate) unpack8(*reinterpret_cast<const uint4*>(dst), acc);
        for (int s = 0; s < splits; ++s) {
            const float4 a = *reinterpret_cast<const float4*>(src + s * slab);
            const float4 b = *reinterpret_cast<const float4*>(src + s * slab + 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        stg16(dst, pack8(acc));
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

template <int BLOCK_N, int kStages, bool A_MN, bool B_MN, int CLUSTER>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, int sms,
           cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N, kStages>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, kStages, A_MN, B_MN, CLUSTER>;
    static bool configured = false;  // per template instantiation
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
        configured = true;
    }
    const int m_units = ((M + BLOCK_M - 1) / BLOCK_M + CLUSTER - 1) / CLUSTER;
    const int units = m_units * ((N + BLOCK_N - 1) / BLOCK_N) * ep.split_k;
    const int max_clusters = sms / CLUSTER;
    const int grid = (units < max_clusters ? units : max_clusters) * CLUSTER;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = L::kDynamic;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // band height: balance distinct A tiles (128 rows each per CTA) against distinct B tiles (BLOCK_N rows) per wave
    int group = 16 / CLUSTER;
    if (group > m_units) group = m_units;
    VLK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, group, ep));
    VLK_CHECK_LAUNCH("vlk_gemm_bf16");
    return VLK_OK;
}

template <int BLOCK_N, int kStages, bool A_MN, bool B_MN>
int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, int sms,
                cudaStream_t stream) {
    using L = SmemLayout2<BLOCK_N, kStages>;
    auto kern = gemm_bf16_2cta_kernel<BLOCK_N, kStages, A_MN, B_MN>;
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
        configured = true;
    }
    const int m_units = ((M + BLOCK_M - 1) / BLOCK_M + 1) / 2;
    const int units = m_units * ((N + BLOCK_N - 1) / BLOCK_N) * ep.split_k;
    const int max_clusters = sms / 2;
    const int grid = (units < max_clusters ? units : max_clusters) * 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = L::kDynamic;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int group = 8;
    if (group > m_units) group = m_units;
    VLK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, M, N, K, group, ep));
    VLK_CHECK_LAUNCH("vlk_gemm_bf16(2cta)");
    return VLK_OK;
}

// Relative cost of finishing the problem with a given BLOCK_N: waves x per-tile time.  A 128xBN tile needs
// max(BN/2 tensor cycles, (128+BN)/4 smem-read cycles) per 16-deep k-step (B300_MICROARCH: floor = M*N/256,
// smem crossbar 128 B/clk).
double tile_cost(int M, int N, int bn, int cluster, int sms) {
    const long m_units = ((M + BLOCK_M - 1) / BLOCK_M + cluster - 1) / cluster;
    const long units = m_units * ((N + bn - 1) / bn);
    const long slots = sms / cluster;
    const long waves = (units + slots - 1) / slots;
    // per k-step of 16: tensor cycles bn/2, smem-read cycles (128+bn)/4, L2 delivery (128 + bn/cluster) rows of
    // 32 B at ~45 B/clk/SM
    const double per_tile = fmax(fmax(bn / 2.0, (128.0 + bn) / 4.0), (128.0 + bn / cluster) * 32.0 / 45.0) + 6.0;
    return waves * per_tile;
}

template <bool A_MN, bool B_MN>
int dispatch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const EpiParams& ep, int sms, int bn,
             int cluster, cudaStream_t stream) {
    if (cluster == 3) {  // cta_group::2 pair
        if (bn == 256) return launch_2cta<256, 6, A_MN, B_MN>(ta, tb, M, N, K, ep, sms, stream);
        return launch_2cta<128, 8, A_MN, B_MN>(ta, tb, M, N, K, ep, sms, stream);
    }
#ifdef VLK_GEMM_MULTICAST_VARIANT   // single-CTA MMA with TMA multicast of B across a CTA pair: measured, never the
    if (cluster == 2) {             // fastest (profiles/r01_gemm_sweep2.log); compiled only for that comparison
        if (bn == 256) return launch<256, 3, A_MN, B_MN, 2>(ta, tb, M, N, K, ep, sms, stream);
        return launch<128, 5, A_MN, B_MN, 2>(ta, tb, M, N, K, ep, sms, stream);
    }
#else
    if (cluster == 2) cluster = 1;
#endif
    switch (bn) {
        case 256:
            return launch<256, 4, A_MN, B_MN, 1>(ta, tb, M, N, K, ep, sms, stream);
        case 128:
            return launch<128, 6, A_MN, B_MN, 1>(ta, tb, M, N, K, ep, sms, stream);
        default:
            return launch<64, 8, A_MN, B_MN, 1>(ta, tb, M, N, K, ep, sms, stream);
    }
}

}  // namespace

int splitk_reduce(const float* ws, int splits, long long slab, void* out, int M, int N, int ldd, int accumulate,
                  cudaStream_t stream) {
    const long long work = slab / 8;
    const int sms = device_sm_count();
    long long blocks = (work + 255) / 256;
    if (blocks > sms * 8LL) blocks = sms * 8LL;
    splitk_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(ws, splits, slab, static_cast<bf16*>(out), M, N,
                                                                            ldd, accumulate);
    VLK_CHECK_LAUNCH("vlk_gemm_bf16_splitk(reduce)");
    return VLK_OK;
}
}  // namespace vlk

using namespace vlk;

int vlk::gemm_tile_n(int N) { return N >= 192 ? 256 : (N >= 96 ? 128 : 64); }

int vlk::gemm_impl(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                   int transA, int transB, const void* bias, const void* residual, int ldr, const void* aux_in,
                   void* aux_out, int ld_aux, const float* scale, int act, int dact, float alpha, int out_fp32,
                   int split_k, long long split_stride, int* split_used, void* stream,
                   const float* ln_mean, const float* ln_rstd, const float* ln_colsum,
                   const float* ln_sums, float ln_eps, float* stats_out,
                   int bn_override, int pair_override, const CeEpilogue* ce) {
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
    if (split_k < 1) split_k = 1;
    VLK_REQUIRE(split_k == 1 || (out_fp32 && !bias && !residual && !aux_in && !aux_out && !scale && act == VLK_ACT_NONE),
                VLK_ERR_INVALID_ARG,
                "vlk_gemm_bf16: split_k > 1 accumulates raw partial products atomically: it needs out_fp32 (zeroed by "
                "the caller) and no epilogue operands");
    {
        // no empty slices: an empty slice would add an un-initialised accumulator
        const int kblocks = (K + BLOCK_K - 1) / BLOCK_K;
        if (split_k > kblocks) split_k = kblocks;
        const int per = (kblocks + split_k - 1) / split_k;
        split_k = (kblocks + per - 1) / per;
    }
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
    ep.split_k = split_k;
    ep.split_stride = split_k > 1 ? split_stride : 0;
    if (split_used) *split_used = split_k;
    ep.ln_mean = ln_mean;
    ep.ln_rstd = ln_rstd;
    ep.ln_colsum = ln_colsum;
    ep.ln_sums = ln_sums;
    ep.ln_inv_k = 1.0f / static_cast<float>(K);
    ep.ln_eps = ln_eps;
    ep.stats_out = stats_out;
    ep.ce_mode = 0;
    ep.ce_labels = nullptr;
    ep.ce_partial = nullptr;
    ep.ce_label_logit = nullptr;
    ep.ce_lse = nullptr;
    ep.ce_row_scale = nullptr;
    ep.ce_col0 = 0;
    {
        auto wide = [](const void* p, int ld) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 31u) == 0 && ld % 16 == 0); };
        ep.wide_io = wide(D, ldd) && wide(residual, ldr) && wide(aux_in, ld_aux) && wide(aux_out, ld_aux);
    }
    if (ce != nullptr) {
        ep.ce_mode = ce->mode;
        ep.ce_labels = ce->labels;
        ep.ce_partial = reinterpret_cast<float2*>(ce->partial);
        ep.ce_label_logit = ce->label_logit;
        ep.ce_lse = ce->lse;
        ep.ce_row_scale = ce->row_scale;
        ep.ce_col0 = ce->col0;
    }
#ifdef VLK_BRINGUP
    ep.debug = 0;
    if (const char* f = getenv("VLK_GEMM_DEBUG")) ep.debug = atoi(f);
    ep.dbg_out = g_gemm_dbg_out;
#endif

    // Tile selection (measured on B200, profiles/r01_gemm_sweep*.log): the 256-wide tile wins on every shape of
    // this path (fewer, longer tiles amortise the epilogue; N = 768/1024/2304/3072/4096/50304 all tile well), and
    // the cta_group::2 pair wins or ties whenever there are at least two 128-row blocks.
    int bn = gemm_tile_n(N);
    int cluster = (M > BLOCK_M && bn >= 128 && sms % 2 == 0) ? 3 : 1;
    (void)tile_cost;
    if (bn_override == 256 || bn_override == 128 || bn_override == 64) {   // vlk_gemm_bf16_tile only
        bn = bn_override;
        if (bn == 64) cluster = 1;
    }
    if (pair_override == 0) cluster = 1;
    else if (pair_override == 1 && bn >= 128 && sms % 2 == 0) cluster = 3;
#ifndef VLK_GEMM_MULTICAST_VARIANT
    if (cluster == 2) cluster = 1;   // the multicast pair is not compiled in (see dispatch)
#endif

    CUtensorMap ta, tb;
    int rc;
    if (!transA)
        rc = make_tmap(&ta, A, K, M, lda, BLOCK_K, BLOCK_M);
    else
        rc = make_tmap(&ta, A, M, K, lda, 64, BLOCK_K);
    if (rc) return rc;
    if (!transB)
        rc = make_tmap(&tb, B, K, N, ldb, BLOCK_K, bn / (cluster >= 2 ? 2 : 1));  // one box = this CTA's slice of B
    else
        rc = make_tmap(&tb, B, N, K, ldb, 64, BLOCK_K);
    if (rc) return rc;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!transA && !transB) return dispatch<false, false>(ta, tb, M, N, K, ep, sms, bn, cluster, s);
    if (!transA && transB) return dispatch<false, true>(ta, tb, M, N, K, ep, sms, bn, cluster, s);
    if (transA && !transB) return dispatch<true, false>(ta, tb, M, N, K, ep, sms, bn, cluster, s);
    return dispatch<true, true>(ta, tb, M, N, K, ep, sms, bn, cluster, s);
}

#ifdef VLK_BRINGUP
extern "C" void vlk_bringup_set_gemm_debug(void* p) { g_gemm_dbg_out = static_cast<unsigned long long*>(p); }
#endif

extern "C" int vlk_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                             int transA, int transB, const void* bias, const void* residual, int ldr,
                             const void* aux_in, void* aux_out, int ld_aux, const float* scale, int act, int dact,
                             float alpha, int out_fp32, int split_k, void* stream) {
    return gemm_impl(A, B, D, M, N, K, lda, ldb, ldd, transA, transB, bias, residual, ldr, aux_in, aux_out, ld_aux, scale,
                     act, dact, alpha, out_fp32, split_k, 0, nullptr, stream);
}

extern "C" int vlk_gemm_bf16_tile(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                                  int transA, int transB, const void* bias, const void* residual, int ldr, int act,
                                  int tile_n, int cta_pair, void* stream) {
    VLK_REQUIRE(tile_n == 0 || tile_n == 64 || tile_n == 128 || tile_n == 256, VLK_ERR_INVALID_ARG,
                "vlk_gemm_bf16_tile: tile_n=%d (0 = automatic, 64, 128, 256)", tile_n);
    return gemm_impl(A, B, D, M, N, K, lda, ldb, ldd, transA, transB, bias, residual, ldr, nullptr, nullptr, 0, nullptr, act,
                     0, 1.0f, 0, 1, 0, nullptr, stream, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, tile_n, cta_pair);
}

extern "C" int vlk_gemm_bf16_splitk(const void* A, const void* B, void* D, float* workspace, int M, int N, int K,
                                    int lda, int ldb, int ldd, int transA, int transB, float alpha, int split_k,
                                    int accumulate, void* stream) {
    VLK_REQUIRE(workspace != nullptr && aligned16(workspace), VLK_ERR_INVALID_ARG,
                "vlk_gemm_bf16_splitk: needs a 16B-aligned fp32 workspace of split_k * M * N elements");
    VLK_REQUIRE(split_k >= 1 && split_k <= 64, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16_splitk: split_k=%d", split_k);
    VLK_REQUIRE(D && aligned16(D) && ldd % 8 == 0, VLK_ERR_ALIGNMENT, "vlk_gemm_bf16_splitk: D");
    int used = 1;
    const long long slab = static_cast<long long>(M) * N;
    // slab mode needs split_k >= 2 inside gemm_impl; a single slice is just an fp32 GEMM into slab 0
    int rc = gemm_impl(A, B, workspace, M, N, K, lda, ldb, N, transA, transB, nullptr, nullptr, 0, nullptr, nullptr, 0,
                       nullptr, VLK_ACT_NONE, 0, alpha, 1, split_k, slab, &used, stream);
    if (rc) return rc;
    return splitk_reduce(workspace, used, slab, D, M, N, ldd, accumulate, static_cast<cudaStream_t>(stream));
}

extern "C" int vlk_gemm_bf16_lnfold(const void* X, const void* Wf, void* D, int M, int N, int K, int ldx, int ldw, int ldd,
                                    const void* bias, const float* row_mean, const float* row_rstd,
                                    const float* col_sum, int act, void* stream) {
    VLK_REQUIRE(row_mean && row_rstd && col_sum, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16_lnfold: null statistics");
    VLK_REQUIRE(aligned16(col_sum), VLK_ERR_ALIGNMENT, "vlk_gemm_bf16_lnfold: col_sum must be 16-byte aligned");
    VLK_REQUIRE(act == VLK_ACT_NONE || act == VLK_ACT_QUICK_GELU || act == VLK_ACT_GELU_TANH, VLK_ERR_UNSUPPORTED,
                "vlk_gemm_bf16_lnfold: act=%d", act);
    return gemm_impl(X, Wf, D, M, N, K, ldx, ldw, ldd, 0, 0, bias, nullptr, 0, nullptr, nullptr, 0, nullptr, act, 0, 1.0f,
                     0, 1, 0, nullptr, stream, row_mean, row_rstd, col_sum);
}

extern "C" int vlk_gemm_bf16_lnfold_sums(const void* X, const void* Wf, void* D, int M, int N, int K, int ldx, int ldw,
                                         int ldd, const void* bias, const float* row_sums, float eps,
                                         const float* col_sum, int act, void* stream) {
    VLK_REQUIRE(row_sums && col_sum, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16_lnfold_sums: null statistics");
    VLK_REQUIRE(aligned16(col_sum) && (reinterpret_cast<uintptr_t>(row_sums) & 7u) == 0, VLK_ERR_ALIGNMENT,
                "vlk_gemm_bf16_lnfold_sums: col_sum must be 16-byte, row_sums 8-byte aligned");
    VLK_REQUIRE(act == VLK_ACT_NONE || act == VLK_ACT_QUICK_GELU || act == VLK_ACT_GELU_TANH, VLK_ERR_UNSUPPORTED,
                "vlk_gemm_bf16_lnfold_sums: act=%d", act);
    return gemm_impl(X, Wf, D, M, N, K, ldx, ldw, ldd, 0, 0, bias, nullptr, 0, nullptr, nullptr, 0, nullptr, act, 0, 1.0f,
                     0, 1, 0, nullptr, stream, nullptr, nullptr, col_sum, row_sums, eps, nullptr);
}

extern "C" int vlk_gemm_bf16_stats(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                                   const void* bias, const void* residual, int ldr, float* stats_out, void* stream) {
    VLK_REQUIRE(stats_out != nullptr, VLK_ERR_INVALID_ARG, "vlk_gemm_bf16_stats: null stats_out");
    VLK_REQUIRE(N >= 96, VLK_ERR_UNSUPPORTED, "vlk_gemm_bf16_stats: N=%d (needs the staged epilogue, N >= 96)", N);
    return gemm_impl(A, B, D, M, N, K, lda, ldb, ldd, 0, 0, bias, residual, ldr, nullptr, nullptr, 0, nullptr, VLK_ACT_NONE,
                     0, 1.0f, 0, 1, 0, nullptr, stream, nullptr, nullptr, nullptr, nullptr, 0.f, stats_out);
}
