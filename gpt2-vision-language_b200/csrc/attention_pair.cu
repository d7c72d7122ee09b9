// Short-sequence attention (Tq, Tk <= 64, head dim 64) on the tcgen05 tensor cores, forward and backward.
//
// This is the shape of every attention of the captioning step outside CLIP: GPT-2 over [33 image + 31 text] = 64
// tokens (63 with the Q-Former prefix), the gated cross-attention 31 x 33, the Q-Former's 32 x 32 / 32 x 33.  A single
// head only fills 64 of the 128 rows of a tcgen05.mma, so one CTA takes TWO (batch, head) problems and stacks them:
//
//     S = [Q_a ; Q_b] [K_a ; K_b]^T   (128 x 128; the two diagonal 64 x 64 blocks are the two heads' scores)
//     O = P [V_a ; V_b]               (P = softmax of the diagonal blocks, ZERO in the off-diagonal blocks)
//
// so every product is one M = 128 MMA chain and no operand is ever reshaped; the wasted off-diagonal FLOPs are free
// at this size (the kernel is bounded by its ~1.5 us TMA -> MMA -> softmax -> MMA -> store latency chain, not by
// the tensor pipe).  128 threads, thread = TMEM lane = one row of one of the two heads.
//
// Backward, one kernel, no atomics: phase A computes everything TRANSPOSED (thread = key row) so that the operand
// coming from TMEM is always the A operand — S^T = K Q^T, dP^T = V dO^T, P^T / dS^T written in place as bf16,
// dV = P^T dO, dK = dS^T Q — and phase B recomputes S = Q K^T, dP = dO V^T with thread = query row for
// dQ = dS K.  The four smem tiles (Q, K, V, dO; 128-byte swizzle) serve as K-major operands of the score products
// and as MN-major B operands of the gradient products.
//
// Attention-probability dropout (nn.MultiheadAttention(dropout=0.1) in the Q-Former, gpt2_q_former/model.py:119,123)
// uses the same Philox counters as the CUDA-core kernel in attention_small.cu: element ((b*H + h) * 64 + i) * 64 + j.
#include <cuda.h>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace vlk {
namespace {

constexpr int kHalfTile = 64 * 128;   // bytes of one head's 64-row x 64-col bf16 tile
constexpr int kPairTile = 2 * kHalfTile;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Strides {
    long long bs;
    int rs;
};

struct PairParams {
    // forward: o, lse; backward: o, d_o (for delta), lse, dq, dk, dv
    bf16* o;
    const bf16* o_in;
    const bf16* d_o;
    float* lse;
    const float* lse_in;
    bf16 *dq, *dk, *dv;
    Strides os, dqs, dks, dvs;
    int B, H, Tq, Tk, causal;
    float scale, scale_log2e, dropout_p;
    const unsigned long long* seed_state;
    uint32_t stream_id;
};

// 64 bf16 results of one row -> global (row-per-thread: 128 contiguous bytes)
__device__ __forceinline__ void store_row64(bf16* dst, const uint32_t (&r0)[32], const uint32_t (&r1)[32], float mul) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __uint_as_float(r0[q * 8 + u]) * mul;
        stg16(dst + q * 8, pack8(t));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __uint_as_float(r1[q * 8 + u]) * mul;
        stg16(dst + 32 + q * 8, pack8(t));
    }
}

__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {  // 32 packed columns (= 64 bf16) of zeros
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    ptx::tmem_st_32x32b_x16(taddr, z);
    ptx::tmem_st_32x32b_x16(taddr + 16, z);
}

// ================================================================================================
// forward.  smem: Q | K | V (16 KB each: head a rows 0..63, head b rows 64..127) | barriers.  TMEM: 128 columns.
// ================================================================================================
__global__ void __launch_bounds__(128, 4)
attn_pair_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, PairParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kPairTile;
    uint8_t* sV = smem + 2 * kPairTile;
    uint64_t* bar_in = reinterpret_cast<uint64_t*>(smem + 3 * kPairTile);
    uint64_t* bar_s = bar_in + 1;
    uint64_t* bar_o = bar_in + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_in + 3);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int hh = warp >> 1;                      // which of the two stacked heads this thread's row belongs to
    const int i = threadIdx.x & 63;                // query row inside the head
    const int unit = 2 * blockIdx.x + hh;          // flattened (b, h)
    const bool valid = unit < p.B * p.H;
    const int b = valid ? unit / p.H : p.B, h = valid ? unit % p.H : 0;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        ptx::mbar_init(bar_in, 1);
        ptx::mbar_init(bar_s, 1);
        ptx::mbar_init(bar_o, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 128);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tS = tmem, tO = tmem + 64;

    if (warp == 0) {   // converged warp, one elected lane issues (MMA operands in uniform registers)
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(bar_in, 3 * kPairTile);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int u = 2 * blockIdx.x + t;
                const bool ok = u < p.B * p.H;
                const int bb = ok ? u / p.H : p.B, hx = ok ? u % p.H : 0;  // batch index B is out of bounds: zero fill
                ptx::tma_load_3d(sQ + t * kHalfTile, &tmap_q, bar_in, hx * 64, 0, bb);
                ptx::tma_load_3d(sK + t * kHalfTile, &tmap_k, bar_in, hx * 64, 0, bb);
                ptx::tma_load_3d(sV + t * kHalfTile, &tmap_v, bar_in, hx * 64, 0, bb);
            }
        }
        __syncwarp();
        ptx::mbar_wait(bar_in, 0);
        ptx::tc_fence_after_sync();
        const uint32_t aq = ptx::smem_u32(sQ), ak = ptx::smem_u32(sK);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128, 0, 0);
        if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ptx::umma_bf16_ss(tS, ptx::make_smem_desc_sw128(aq + k * 32, 16, 1024),
                                  ptx::make_smem_desc_sw128(ak + k * 32, 16, 1024), idesc, k != 0);
            ptx::umma_commit(bar_s);
        }
        __syncwarp();
    }
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int shift = p.Tk - p.Tq;
    const int lim = p.causal ? min(p.Tk, i + shift + 1) : p.Tk;   // keys [0, lim) are visible to this row

    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after_sync();
    uint32_t s0[32], s1[32];
    ptx::tmem_ld_32x32b_x32(tS + lane_base + hh * 64, s0);
    ptx::tmem_ld_32x32b_x32(tS + lane_base + hh * 64 + 32, s1);
    ptx::tmem_ld_wait();
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < lim) m = fmaxf(m, __uint_as_float(s0[j]));
        if (j + 32 < lim) m = fmaxf(m, __uint_as_float(s1[j]));
    }
    const float mb = (m == -INFINITY) ? 0.f : m * p.scale_log2e;
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float e0 = j < lim ? ex2f(fmaf(__uint_as_float(s0[j]), p.scale_log2e, -mb)) : 0.f;
        const float e1 = j + 32 < lim ? ex2f(fmaf(__uint_as_float(s1[j]), p.scale_log2e, -mb)) : 0.f;
        l += e0 + e1;
        s0[j] = __float_as_uint(e0);
        s1[j] = __float_as_uint(e1);
    }
    if (p.dropout_p > 0.f) {  // dropout on the probabilities; the normaliser l is the un-dropped sum
        const DropoutKey dk = make_dropout_key(p.seed_state, p.stream_id, p.dropout_p);
        const unsigned long long g0 = (static_cast<unsigned long long>(unit) * 64 + i) * 16;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float mk0[4], mk1[4];
            dropout_scales4(dk, g0 + g, mk0);
            dropout_scales4(dk, g0 + 8 + g, mk1);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s0[4 * g + u] = __float_as_uint(__uint_as_float(s0[4 * g + u]) * mk0[u]);
                s1[4 * g + u] = __float_as_uint(__uint_as_float(s1[4 * g + u]) * mk1[u]);
            }
        }
    }
    {
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const bf162 h2 = __floats2bfloat162_rn(__uint_as_float(s0[2 * t]), __uint_as_float(s0[2 * t + 1]));
            pk[t] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        ptx::tmem_st_32x32b_x16(tS + lane_base + hh * 32, pk);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const bf162 h2 = __floats2bfloat162_rn(__uint_as_float(s1[2 * t]), __uint_as_float(s1[2 * t + 1]));
            pk[t] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        ptx::tmem_st_32x32b_x16(tS + lane_base + hh * 32 + 16, pk);
        tmem_zero32(tS + lane_base + (1 - hh) * 32);   // the other head's keys contribute nothing to this row
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {   // converged warp, one elected lane issues (MMA operands in uniform registers)
        ptx::tc_fence_after_sync();
        const uint32_t av = ptx::smem_u32(sV);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);   // V [key x 64] read MN-major
        if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                ptx::umma_bf16_ts(tO, tS + k * 8, ptx::make_smem_desc_sw128(av + k * 2048, 8192, 1024), idesc, k != 0);
            ptx::umma_commit(bar_o);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after_sync();
    ptx::tmem_ld_32x32b_x32(tO + lane_base, s0);
    ptx::tmem_ld_32x32b_x32(tO + lane_base + 32, s1);
    ptx::tmem_ld_wait();
    if (valid && i < p.Tq) {
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        store_row64(p.o + b * p.os.bs + static_cast<size_t>(i) * p.os.rs + h * 64, s0, s1, inv);
        if (p.lse != nullptr) p.lse[static_cast<size_t>(unit) * p.Tq + i] = m * p.scale + __logf(l);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, 128);
    }
}

// ================================================================================================
// backward.  smem: Q | K | V | dO (16 KB each) | lse*log2e [128] | delta [128] | barriers.  TMEM: 256 columns:
//   R0 = [0,128): S^T -> P^T (bf16, cols 0..63) + dV (cols 64..127);  then S -> dQ (cols 0..63)
//   R1 = [128,256): dP^T -> dS^T + dK;  then dP -> dS
// ================================================================================================
__global__ void __launch_bounds__(128, 2)
attn_pair_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     PairParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kPairTile;
    uint8_t* sV = smem + 2 * kPairTile;
    uint8_t* sdO = smem + 3 * kPairTile;
    float* sLse = reinterpret_cast<float*>(smem + 4 * kPairTile);
    float* sDelta = sLse + 128;
    uint64_t* bar_in = reinterpret_cast<uint64_t*>(sDelta + 128);
    uint64_t* bar_s = bar_in + 1;
    uint64_t* bar_acc = bar_in + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_in + 3);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int hh = warp >> 1;
    const int i = threadIdx.x & 63;               // this thread's row inside its head: key row in phase A, query row in B
    const int unit = 2 * blockIdx.x + hh;
    const bool valid = unit < p.B * p.H;
    const int b = valid ? unit / p.H : p.B, h = valid ? unit % p.H : 0;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        ptx::prefetch_tensormap(&tmap_do);
        ptx::mbar_init(bar_in, 1);
        ptx::mbar_init(bar_s, 1);
        ptx::mbar_init(bar_acc, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tR0 = tmem, tR1 = tmem + 128;

    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(bar_in, 4 * kPairTile);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int u = 2 * blockIdx.x + t;
            const bool ok = u < p.B * p.H;
            const int bb = ok ? u / p.H : p.B, hx = ok ? u % p.H : 0;
            ptx::tma_load_3d(sQ + t * kHalfTile, &tmap_q, bar_in, hx * 64, 0, bb);
            ptx::tma_load_3d(sK + t * kHalfTile, &tmap_k, bar_in, hx * 64, 0, bb);
            ptx::tma_load_3d(sV + t * kHalfTile, &tmap_v, bar_in, hx * 64, 0, bb);
            ptx::tma_load_3d(sdO + t * kHalfTile, &tmap_do, bar_in, hx * 64, 0, bb);
        }
    }
    // row statistics of query row i of this thread's head: delta_i = dO_i . O_i and lse_i (pre-multiplied by log2 e)
    {
        float d = 0.f, ls = INFINITY;   // padded rows: lse = +inf makes every probability of that row 0
        if (valid && i < p.Tq) {
            const size_t off = b * p.os.bs + static_cast<size_t>(i) * p.os.rs + h * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float a[8], g[8];
                unpack8(ldg16(p.o_in + off + c * 8), a);
                unpack8(ldg16(p.d_o + off + c * 8), g);
#pragma unroll
                for (int u = 0; u < 8; ++u) d = fmaf(a[u], g[u], d);
            }
            ls = p.lse_in[static_cast<size_t>(unit) * p.Tq + i] * kLog2e;
        }
        sLse[threadIdx.x] = ls;
        sDelta[threadIdx.x] = d;
    }
    if (warp == 0) {   // converged warp, one elected lane issues (MMA operands in uniform registers)
        ptx::mbar_wait(bar_in, 0);
        ptx::tc_fence_after_sync();
        const uint32_t aq = ptx::smem_u32(sQ), ak = ptx::smem_u32(sK), av = ptx::smem_u32(sV), ado = ptx::smem_u32(sdO);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128, 0, 0);
        if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // S^T = K Q^T ; dP^T = V dO^T
                ptx::umma_bf16_ss(tR0, ptx::make_smem_desc_sw128(ak + k * 32, 16, 1024),
                                  ptx::make_smem_desc_sw128(aq + k * 32, 16, 1024), idesc, k != 0);
                ptx::umma_bf16_ss(tR1, ptx::make_smem_desc_sw128(av + k * 32, 16, 1024),
                                  ptx::make_smem_desc_sw128(ado + k * 32, 16, 1024), idesc, k != 0);
            }
            ptx::umma_commit(bar_s);
        }
        __syncwarp();
    }
    __syncthreads();  // sLse / sDelta visible
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int shift = p.Tk - p.Tq;
    const bool drop = p.dropout_p > 0.f;
    DropoutKey dkey;
    if (drop) dkey = make_dropout_key(p.seed_state, p.stream_id, p.dropout_p);
    const unsigned long long elem0 = static_cast<unsigned long long>(unit) * 64 * 64;
    const float* stL = sLse + hh * 64;
    const float* stD = sDelta + hh * 64;

    // ---------------- phase A: thread = key row kj = i ----------------
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after_sync();
    {
        const int kj = i;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
            uint32_t rs[32], rp[32];
            ptx::tmem_ld_32x32b_x32(tR0 + lane_base + hh * 64 + c, rs);
            ptx::tmem_ld_32x32b_x32(tR1 + lane_base + hh * 64 + c, rp);
            ptx::tmem_ld_wait();
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                float pv[2], dv[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int qi = c + 2 * t + u;  // query row of this head
                    float pr = ex2f(fmaf(__uint_as_float(rs[2 * t + u]), p.scale_log2e, -stL[qi]));
                    if (kj >= p.Tk || (p.causal && kj > qi + shift)) pr = 0.f;
                    float mk = 1.f;
                    if (drop) mk = dropout_scale1(dkey, elem0 + static_cast<unsigned long long>(qi) * 64 + kj);
                    pv[u] = pr * mk;                                                // dropped probabilities feed dV
                    dv[u] = pr * (__uint_as_float(rp[2 * t + u]) * mk - stD[qi]);    // dS = P o (mask o dP - delta)
                }
                const bf162 hp = __floats2bfloat162_rn(pv[0], pv[1]);
                const bf162 hd = __floats2bfloat162_rn(dv[0], dv[1]);
                pk[t] = *reinterpret_cast<const uint32_t*>(&hp);
                dk[t] = *reinterpret_cast<const uint32_t*>(&hd);
            }
            ptx::tmem_st_32x32b_x16(tR0 + lane_base + hh * 32 + (c >> 1), pk);   // P^T in place
            ptx::tmem_st_32x32b_x16(tR1 + lane_base + hh * 32 + (c >> 1), dk);   // dS^T in place
        }
        tmem_zero32(tR0 + lane_base + (1 - hh) * 32);
        tmem_zero32(tR1 + lane_base + (1 - hh) * 32);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {   // converged warp, one elected lane issues (MMA operands in uniform registers)
        ptx::tc_fence_after_sync();
        const uint32_t bq = ptx::smem_u32(sQ), bdo = ptx::smem_u32(sdO);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);  // B = [query x 64] read MN-major
        if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                ptx::umma_bf16_ts(tR0 + 64, tR0 + k * 8, ptx::make_smem_desc_sw128(bdo + k * 2048, 8192, 1024), idesc, k != 0);
                ptx::umma_bf16_ts(tR1 + 64, tR1 + k * 8, ptx::make_smem_desc_sw128(bq + k * 2048, 8192, 1024), idesc, k != 0);
            }
            ptx::umma_commit(bar_acc);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar_acc, 0);
    ptx::tc_fence_after_sync();
    {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32b_x32(tR0 + 64 + lane_base, r0);
        ptx::tmem_ld_32x32b_x32(tR0 + 96 + lane_base, r1);
        ptx::tmem_ld_wait();
        if (valid && i < p.Tk)
            store_row64(p.dv + b * p.dvs.bs + static_cast<size_t>(i) * p.dvs.rs + h * 64, r0, r1, 1.0f);
        ptx::tmem_ld_32x32b_x32(tR1 + 64 + lane_base, r0);
        ptx::tmem_ld_32x32b_x32(tR1 + 96 + lane_base, r1);
        ptx::tmem_ld_wait();
        if (valid && i < p.Tk)
            store_row64(p.dk + b * p.dks.bs + static_cast<size_t>(i) * p.dks.rs + h * 64, r0, r1, p.scale);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();  // every TMEM read of phase A is done before phase B's MMAs overwrite R0 / R1

    // ---------------- phase B: thread = query row qi = i ----------------
    if (warp == 0) {   // converged warp, one elected lane issues (MMA operands in uniform registers)
        ptx::tc_fence_after_sync();
        const uint32_t aq = ptx::smem_u32(sQ), ak = ptx::smem_u32(sK), av = ptx::smem_u32(sV), ado = ptx::smem_u32(sdO);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128, 0, 0);
        if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // S = Q K^T ; dP = dO V^T
                ptx::umma_bf16_ss(tR0, ptx::make_smem_desc_sw128(aq + k * 32, 16, 1024),
                                  ptx::make_smem_desc_sw128(ak + k * 32, 16, 1024), idesc, k != 0);
                ptx::umma_bf16_ss(tR1, ptx::make_smem_desc_sw128(ado + k * 32, 16, 1024),
                                  ptx::make_smem_desc_sw128(av + k * 32, 16, 1024), idesc, k != 0);
            }
            ptx::umma_commit(bar_s);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar_s, 1);
    ptx::tc_fence_after_sync();
    {
        const int qi = i;
        const float ls = stL[qi], dl = stD[qi];
        const int lim = p.causal ? min(p.Tk, qi + shift + 1) : p.Tk;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
            uint32_t rs[32], rp[32];
            ptx::tmem_ld_32x32b_x32(tR0 + lane_base + hh * 64 + c, rs);
            ptx::tmem_ld_32x32b_x32(tR1 + lane_base + hh * 64 + c, rp);
            ptx::tmem_ld_wait();
            uint32_t dk[16];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float mk[4] = {1.f, 1.f, 1.f, 1.f};
                if (drop) dropout_scales4(dkey, (elem0 + static_cast<unsigned long long>(qi) * 64 + c) / 4 + g, mk);
                float dv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int kj = c + 4 * g + u;
                    float pr = ex2f(fmaf(__uint_as_float(rs[4 * g + u]), p.scale_log2e, -ls));
                    if (kj >= lim) pr = 0.f;
                    dv[u] = pr * (__uint_as_float(rp[4 * g + u]) * mk[u] - dl);
                }
                const bf162 h0 = __floats2bfloat162_rn(dv[0], dv[1]);
                const bf162 h1 = __floats2bfloat162_rn(dv[2], dv[3]);
                dk[2 * g] = *reinterpret_cast<const uint32_t*>(&h0);
                dk[2 * g + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            }
            ptx::tmem_st_32x32b_x16(tR1 + lane_base + hh * 32 + (c >> 1), dk);   // dS in place
        }
        tmem_zero32(tR1 + lane_base + (1 - hh) * 32);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {   // converged warp, one elected lane issues (MMA operands in uniform registers)
        ptx::tc_fence_after_sync();
        const uint32_t bk = ptx::smem_u32(sK);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);  // K [key x 64] read MN-major
        if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                ptx::umma_bf16_ts(tR0, tR1 + k * 8, ptx::make_smem_desc_sw128(bk + k * 2048, 8192, 1024), idesc, k != 0);
            ptx::umma_commit(bar_acc);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar_acc, 1);
    ptx::tc_fence_after_sync();
    {
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32b_x32(tR0 + lane_base, r0);
        ptx::tmem_ld_32x32b_x32(tR0 + 32 + lane_base, r1);
        ptx::tmem_ld_wait();
        if (valid && i < p.Tq)
            store_row64(p.dq + b * p.dqs.bs + static_cast<size_t>(i) * p.dqs.rs + h * 64, r0, r1, p.scale);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn3() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// [B, T, W] bf16 view -> box 64 cols x 64 rows x 1, 128B swizzle; rows >= T and batch index B read as zeros
int tmap_rows64(CUtensorMap* map, const void* base, int W, int T, int B, int rs, long long bs) {
    EncodeTiledFn fn = encode_fn3();
    VLK_REQUIRE(fn != nullptr, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
    cuuint32_t box[3] = {64, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLK_REQUIRE(r == CUDA_SUCCESS, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled(pair) failed with CUresult %d", (int)r);
    return VLK_OK;
}

constexpr int kFwdSmem = 3 * kPairTile + 64 + 1024;
constexpr int kBwdSmem = 4 * kPairTile + 2 * 128 * 4 + 64 + 1024;

}  // namespace

// Tq, Tk <= 64 and TMA-compatible strides (16-byte multiples; the batch stride only matters when B > 1).
bool attn_pair_applicable(int B, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs,
                          int v_rs, long long o_bs, int o_rs) {
    if (Tq > 64 || Tk > 64) return false;
    if (q_rs % 8 || k_rs % 8 || v_rs % 8 || o_rs % 8) return false;
    if (B > 1 && (q_bs % 8 || k_bs % 8 || v_bs % 8 || o_bs % 8)) return false;
    return true;
}

int attn_pair_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                  long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                  int o_rs, int causal, float scale, float dropout_p, const unsigned long long* seed_state,
                  unsigned int stream_id, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(attn_pair_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
        configured = true;
    }
    CUtensorMap tq, tk, tv;
    int rc = tmap_rows64(&tq, q, H * 64, Tq, B, q_rs, B > 1 ? q_bs : static_cast<long long>(Tq) * q_rs);
    if (rc) return rc;
    rc = tmap_rows64(&tk, k, H * 64, Tk, B, k_rs, B > 1 ? k_bs : static_cast<long long>(Tk) * k_rs);
    if (rc) return rc;
    rc = tmap_rows64(&tv, v, H * 64, Tk, B, v_rs, B > 1 ? v_bs : static_cast<long long>(Tk) * v_rs);
    if (rc) return rc;
    PairParams p = {};
    p.o = static_cast<bf16*>(o);
    p.lse = lse;
    p.os = Strides{o_bs, o_rs};
    p.B = B;
    p.H = H;
    p.Tq = Tq;
    p.Tk = Tk;
    p.causal = causal;
    p.scale = scale;
    p.scale_log2e = scale * kLog2e;
    p.dropout_p = dropout_p;
    p.seed_state = seed_state;
    p.stream_id = stream_id;
    attn_pair_fwd_kernel<<<(B * H + 1) / 2, 128, kFwdSmem, stream>>>(tq, tk, tv, p);
    VLK_CHECK_LAUNCH("vlk_attn_fwd(pair)");
    return VLK_OK;
}

int attn_pair_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                  void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                  int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                  long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float dropout_p,
                  const unsigned long long* seed_state, unsigned int stream_id, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(attn_pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
        configured = true;
    }
    CUtensorMap tq, tk, tv, tdo;
    int rc = tmap_rows64(&tq, q, H * 64, Tq, B, q_rs, B > 1 ? q_bs : static_cast<long long>(Tq) * q_rs);
    if (rc) return rc;
    rc = tmap_rows64(&tk, k, H * 64, Tk, B, k_rs, B > 1 ? k_bs : static_cast<long long>(Tk) * k_rs);
    if (rc) return rc;
    rc = tmap_rows64(&tv, v, H * 64, Tk, B, v_rs, B > 1 ? v_bs : static_cast<long long>(Tk) * v_rs);
    if (rc) return rc;
    rc = tmap_rows64(&tdo, d_o, H * 64, Tq, B, o_rs, B > 1 ? o_bs : static_cast<long long>(Tq) * o_rs);
    if (rc) return rc;
    PairParams p = {};
    p.o_in = static_cast<const bf16*>(o);
    p.d_o = static_cast<const bf16*>(d_o);
    p.lse_in = lse;
    p.dq = static_cast<bf16*>(dq);
    p.dk = static_cast<bf16*>(dk);
    p.dv = static_cast<bf16*>(dv);
    p.os = Strides{o_bs, o_rs};
    p.dqs = Strides{dq_bs, dq_rs};
    p.dks = Strides{dk_bs, dk_rs};
    p.dvs = Strides{dv_bs, dv_rs};
    p.B = B;
    p.H = H;
    p.Tq = Tq;
    p.Tk = Tk;
    p.causal = causal;
    p.scale = scale;
    p.scale_log2e = scale * kLog2e;
    p.dropout_p = dropout_p;
    p.seed_state = seed_state;
    p.stream_id = stream_id;
    attn_pair_bwd_kernel<<<(B * H + 1) / 2, 128, kBwdSmem, stream>>>(tq, tk, tv, tdo, p);
    VLK_CHECK_LAUNCH("vlk_attn_bwd(pair)");
    return VLK_OK;
}

}  // namespace vlk
