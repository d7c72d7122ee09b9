// Attention forward + backward for short sequences (Tq, Tk <= 64, head dim 64): one CTA per (batch, head) keeps
// every operand of that head in shared memory and runs the five 64x64x64 products of the backward pass (three
// for the forward) as register-tiled fp32 FMAs (16x16 threads, 4x4 outputs each, float4 shared-memory reads).
//
// This is the shape of every attention in the captioning step outside CLIP: GPT-2 over [33 image + 31 text] = 64
// tokens (63 for the Q-Former prefix), cross-attention 31 x 33, Q-Former 32 x 32 and 32 x 33.  The whole problem
// is a few GFLOP per step; what matters is one launch per layer with no re-reading, not tensor-core peak.
#include "common.cuh"

namespace vlk {
namespace {

constexpr int T = 64;   // padded sequence tile
constexpr int D = 64;
constexpr int kThreads = 256;

struct Addr {
    long long bs;
    int rs;
};

// natural layout: dst[r][c] (row-major, 64 x 64 fp32); rows >= valid are zero
__device__ __forceinline__ void load_nat(float* __restrict__ dst, const bf16* __restrict__ src, int rs, int valid,
                                         float mul) {
    for (int v = threadIdx.x; v < T * 8; v += kThreads) {
        const int r = v >> 3, c = (v & 7) * 8;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < valid) unpack8(ldg16(src + static_cast<size_t>(r) * rs + c), f);
        float4* d4 = reinterpret_cast<float4*>(dst + r * D + c);
        d4[0] = make_float4(f[0] * mul, f[1] * mul, f[2] * mul, f[3] * mul);
        d4[1] = make_float4(f[4] * mul, f[5] * mul, f[6] * mul, f[7] * mul);
    }
}
// transposed layout: dst[c][r]
__device__ __forceinline__ void load_tr(float* __restrict__ dst, const bf16* __restrict__ src, int rs, int valid,
                                        float mul) {
    for (int v = threadIdx.x; v < T * 8; v += kThreads) {
        const int r = v & 63, c = (v >> 6) * 8;  // consecutive threads -> consecutive rows: conflict-free stores
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < valid) unpack8(ldg16(src + static_cast<size_t>(r) * rs + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[(c + i) * T + r] = f[i] * mul;
    }
}

// acc[i][j] += sum_k At[k][4*ty + i] * Bt[k][4*tx + j]   (both operands stored contraction-major)
__device__ __forceinline__ void mm_tt(const float* __restrict__ At, const float* __restrict__ Bt, int ty, int tx,
                                      int kmax, float (&acc)[4][4], int kmin = 0) {
#pragma unroll 8
    for (int k = kmin; k < kmax; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(At + k * T + 4 * ty);
        const float4 b = *reinterpret_cast<const float4*>(Bt + k * T + 4 * tx);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

__device__ __forceinline__ void zero(float (&a)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// reduce over the 16 threads that share a row group (same ty): lanes differing in the low 4 bits
__device__ __forceinline__ float rowgroup_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float rowgroup_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void store_rows_bf16(bf16* __restrict__ dst, int rs, const float (&acc)[4][4], int ty,
                                                int tx, int valid_rows, float mul) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 4 * ty + i;
        if (r < valid_rows) {
            bf162* p = reinterpret_cast<bf162*>(dst + static_cast<size_t>(r) * rs + 4 * tx);
            p[0] = __floats2bfloat162_rn(acc[i][0] * mul, acc[i][1] * mul);
            p[1] = __floats2bfloat162_rn(acc[i][2] * mul, acc[i][3] * mul);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// forward: smem = Qt | Kt | V | Pt  (4 x 16 KB)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
attn_small_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      bf16* __restrict__ o, float* __restrict__ lse, int H, int Tq, int Tk, Addr qa, Addr ka,
                      Addr va, Addr oa, int causal, float scale, float dropout_p,
                      const unsigned long long* __restrict__ seed_state, uint32_t stream_id) {
    extern __shared__ __align__(16) float sm[];
    float* Qt = sm;
    float* Kt = sm + T * D;
    float* Vn = sm + 2 * T * D;
    float* Pt = sm + 3 * T * D;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    load_tr(Qt, q + b * qa.bs + h * D, qa.rs, Tq, scale);
    load_tr(Kt, k + b * ka.bs + h * D, ka.rs, Tk, 1.f);
    load_nat(Vn, v + b * va.bs + h * D, va.rs, Tk, 1.f);
    __syncthreads();
    const int shift = Tk - Tq;
    float s[4][4];
    zero(s);
    // causal: a 4x4 tile whose first key lies beyond its last query's horizon is entirely masked
    if (!(causal && 4 * tx > 4 * ty + 3 + shift)) mm_tt(Qt, Kt, ty, tx, D, s);
    float rinv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int qi = 4 * ty + i;
        const int lim = causal ? min(Tk, qi + shift + 1) : Tk;
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * tx + j < lim) m = fmaxf(m, s[i][j]);
        m = rowgroup_max(m);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s[i][j] = (4 * tx + j < lim && m > -INFINITY) ? __expf(s[i][j] - m) : 0.f;
            sum += s[i][j];
        }
        sum = rowgroup_sum(sum);
        rinv[i] = sum > 0.f ? 1.0f / sum : 0.f;
        if (lse != nullptr && tx == 0 && qi < Tq) lse[(static_cast<size_t>(b) * H + h) * Tq + qi] = m + __logf(sum);
    }
    if (dropout_p > 0.f) {  // dropout on the (normalised) probabilities, as nn.MultiheadAttention does
        const DropoutKey dk = make_dropout_key(seed_state, stream_id, dropout_p);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mk[4];
            dropout_scales4(dk, (static_cast<unsigned long long>(blockIdx.x) * T + 4 * ty + i) * (T / 4) + tx, mk);
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] *= mk[j];
        }
    }
    // Pt[j][i]
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(Pt + (4 * tx + j) * T + 4 * ty) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();
    float acc[4][4];
    zero(acc);
    // O[i][d] = sum_j Pt[j][i] * V[j][d]; under a causal mask rows 4ty..4ty+3 see keys < 4ty+4+shift only
    mm_tt(Pt, Vn, ty, tx, causal ? min(Tk, 4 * ty + 4 + shift) : Tk, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] *= rinv[i];
    store_rows_bf16(o + b * oa.bs + h * D, oa.rs, acc, ty, tx, Tq, 1.f);
}

// ---------------------------------------------------------------------------------------------------
// backward: smem = A0 | A1 | A2 | A3 | Qn | Kn | dOn  (7 x 16 KB)
//   phase 1: A0=Qt(scaled) A1=Kt A2=dOt A3=Vt        -> S, dP in registers -> P, dS
//   phase 2: A0=P (natural [i][j]) A1=dS (natural) A2=dSt ([j][i])          -> dV, dK, dQ
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
attn_small_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      const bf16* __restrict__ o, const bf16* __restrict__ d_o, const float* __restrict__ lse,
                      bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv, int H, int Tq, int Tk,
                      Addr qa, Addr ka, Addr va, Addr oa, Addr dqa, Addr dka, Addr dva, int causal, float scale,
                      float dropout_p, const unsigned long long* __restrict__ seed_state, uint32_t stream_id) {
    extern __shared__ __align__(16) float sm[];
    float* A0 = sm;
    float* A1 = sm + T * D;
    float* A2 = sm + 2 * T * D;
    float* A3 = sm + 3 * T * D;
    float* Qn = sm + 4 * T * D;
    float* Kn = sm + 5 * T * D;
    float* dOn = sm + 6 * T * D;
    __shared__ float sDelta[T], sLse[T];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const bf16* qb = q + b * qa.bs + h * D;
    const bf16* kb = k + b * ka.bs + h * D;
    const bf16* vb = v + b * va.bs + h * D;
    const bf16* ob = o + b * oa.bs + h * D;
    const bf16* dob = d_o + b * oa.bs + h * D;
    load_tr(A0, qb, qa.rs, Tq, scale);
    load_tr(A1, kb, ka.rs, Tk, 1.f);
    load_tr(A2, dob, oa.rs, Tq, 1.f);
    load_tr(A3, vb, va.rs, Tk, 1.f);
    load_nat(Qn, qb, qa.rs, Tq, 1.f);
    load_nat(Kn, kb, ka.rs, Tk, 1.f);
    load_nat(dOn, dob, oa.rs, Tq, 1.f);
    // delta_i = dO_i . O_i : 4 threads per row
    {
        const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
        float acc = 0.f;
        if (r < Tq) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float a[8], g[8];
                unpack8(ldg16(ob + static_cast<size_t>(r) * oa.rs + part * 16 + c * 8), a);
                unpack8(ldg16(dob + static_cast<size_t>(r) * oa.rs + part * 16 + c * 8), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(a[i], g[i], acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (part == 0) {
            sDelta[r] = acc;
            sLse[r] = r < Tq ? lse[(static_cast<size_t>(b) * H + h) * Tq + r] : 0.f;
        }
    }
    __syncthreads();
    const int shift = Tk - Tq;
    float s[4][4], dp[4][4];
    zero(s);
    zero(dp);
    if (!(causal && 4 * tx > 4 * ty + 3 + shift)) {  // skip tiles that are entirely masked
        mm_tt(A0, A1, ty, tx, D, s);    // S[i][j] (already scaled)
        mm_tt(A2, A3, ty, tx, D, dp);   // dP[i][j] = dO_i . V_j
    }
    DropoutKey dkey;
    if (dropout_p > 0.f) dkey = make_dropout_key(seed_state, stream_id, dropout_p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int qi = 4 * ty + i;
        const int lim = (qi < Tq) ? (causal ? min(Tk, qi + shift + 1) : Tk) : 0;
        const float l = sLse[qi], dl = sDelta[qi];
        float mk[4] = {1.f, 1.f, 1.f, 1.f};
        if (dropout_p > 0.f)  // the forward's mask, regenerated from the same counters
            dropout_scales4(dkey, (static_cast<unsigned long long>(blockIdx.x) * T + qi) * (T / 4) + tx, mk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float p = (4 * tx + j < lim) ? __expf(s[i][j] - l) : 0.f;
            s[i][j] = p * mk[j];                       // dropped probabilities feed dV
            dp[i][j] = p * (dp[i][j] * mk[j] - dl);    // dS = P o (mask o dP_drop - delta)
        }
    }
    __syncthreads();  // everyone is done reading A0..A3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<float4*>(A0 + (4 * ty + i) * T + 4 * tx) = make_float4(s[i][0], s[i][1], s[i][2], s[i][3]);
        *reinterpret_cast<float4*>(A1 + (4 * ty + i) * T + 4 * tx) =
            make_float4(dp[i][0], dp[i][1], dp[i][2], dp[i][3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(A2 + (4 * tx + j) * T + 4 * ty) =
            make_float4(dp[0][j], dp[1][j], dp[2][j], dp[3][j]);
    __syncthreads();
    float acc[4][4];
    // causal: key rows 4ty..4ty+3 are seen by queries i >= 4ty - shift only; query rows 4ty.. see keys < 4ty+4+shift
    const int i_min = causal ? max(0, 4 * ty - shift) : 0;
    const int j_max = causal ? min(Tk, 4 * ty + 4 + shift) : Tk;
    // dV[j][d] = sum_i P[i][j] dO[i][d]
    zero(acc);
    mm_tt(A0, dOn, ty, tx, Tq, acc, i_min);
    store_rows_bf16(dv + b * dva.bs + h * D, dva.rs, acc, ty, tx, Tk, 1.f);
    // dK[j][d] = scale * sum_i dS[i][j] Q[i][d]
    zero(acc);
    mm_tt(A1, Qn, ty, tx, Tq, acc, i_min);
    store_rows_bf16(dk + b * dka.bs + h * D, dka.rs, acc, ty, tx, Tk, scale);
    // dQ[i][d] = scale * sum_j dSt[j][i] K[j][d]
    zero(acc);
    mm_tt(A2, Kn, ty, tx, j_max, acc);
    store_rows_bf16(dq + b * dqa.bs + h * D, dqa.rs, acc, ty, tx, Tq, scale);
}

}  // namespace

bool attn_small_applicable(int Tq, int Tk) { return Tq <= T && Tk <= T; }

int attn_small_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                   long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                   int o_rs, int causal, float scale, float dropout_p, const unsigned long long* seed_state,
                   unsigned int stream_id, cudaStream_t stream) {
    constexpr int smem = 4 * T * D * sizeof(float);
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    attn_small_fwd_kernel<<<B * H, kThreads, smem, stream>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v), static_cast<bf16*>(o),
        lse, H, Tq, Tk, Addr{q_bs, q_rs}, Addr{k_bs, k_rs}, Addr{v_bs, v_rs}, Addr{o_bs, o_rs}, causal, scale,
        dropout_p, seed_state, stream_id);
    VLK_CHECK_LAUNCH("vlk_attn_fwd(small)");
    return VLK_OK;
}

int attn_small_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                   int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                   long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float dropout_p,
                   const unsigned long long* seed_state, unsigned int stream_id, cudaStream_t stream) {
    constexpr int smem = 7 * T * D * sizeof(float);
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    attn_small_bwd_kernel<<<B * H, kThreads, smem, stream>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
        static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), lse, static_cast<bf16*>(dq), static_cast<bf16*>(dk),
        static_cast<bf16*>(dv), H, Tq, Tk, Addr{q_bs, q_rs}, Addr{k_bs, k_rs}, Addr{v_bs, v_rs}, Addr{o_bs, o_rs},
        Addr{dq_bs, dq_rs}, Addr{dk_bs, dk_rs}, Addr{dv_bs, dv_rs}, causal, scale, dropout_p, seed_state, stream_id);
    VLK_CHECK_LAUNCH("vlk_attn_bwd(small)");
    return VLK_OK;
}

}  // namespace vlk
