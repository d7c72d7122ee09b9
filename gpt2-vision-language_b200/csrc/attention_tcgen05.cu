// Attention forward on the tcgen05 tensor cores for mid-length sequences (64 < Tk <= 272: the CLIP ViT-L/14
// encoder has 257 tokens, 16 heads x 64) — the persistent, warp-specialised kernel `attn_fwd_tcgen05_v4_kernel`:
//
//   TMA:  Q [128 x 64] per query block, K / V [Tk x 64] once per (batch, head), double-buffered across units
//         (128B swizzle, zero-filled past the sequence)
//   MMA1: S = Q K^T            (M=128, N=256; fp32 in TMEM; the 257th key is handled on the CUDA cores)
//   softmax: thread t owns TMEM lane t = one query row -> no cross-thread reduction at all;
//            row max, exp2, row sum in registers; un-normalised bf16 P written IN PLACE over the score columns
//   MMA2: O = P V              (A from TMEM, V as an MN-major B operand, i.e. no transpose)
//   epilogue: O / rowsum -> bf16 -> global; optional log-sum-exp for the backward pass.
//
// The whole key range fits one tile, so there is no online-softmax rescaling here; longer sequences use the
// streaming kernel in attention_flash.cu.  The entry point `attn_mid_fwd` is called by attention_api.cu.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace vlk {

#ifdef VLK_BRINGUP
// bring-up instrumentation (never compiled into the shipped library): per-phase %globaltimer stamps of the first
// tiles, read back by vlk_debug_dump
__device__ long long g_attn_dbg[64 * 16];
#endif

namespace {

constexpr int kQBytes = 128 * 128;                 // 128 rows x 64 bf16


__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// dot of this thread's (swizzled, K-major) Q row with extra-key row e; both 64 bf16
__device__ __forceinline__ float dot_q_kx(const uint8_t* qrow_base, int qr7, const uint8_t* kx, int e) {
    const uint8_t* krow = kx + (e >> 3) * 1024 + (e & 7) * 128;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a[8], b[8];
        unpack8(*reinterpret_cast<const uint4*>(qrow_base + ((j ^ qr7) << 4)), a);
        unpack8(*reinterpret_cast<const uint4*>(krow + ((j ^ (e & 7)) << 4)), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
    }
    return acc;
}
// ------------------------------------------------------------------------------------------------
constexpr int k4KVStage = 68 * 1024;                       // K 32 KB | Kx 2 KB | V 32 KB | Vx 2 KB
constexpr int k4OffQ = 2 * k4KVStage;                      // two Q slots of 16 KB
constexpr int k4OffXchg = k4OffQ + 2 * 16 * 1024;          // per group: extra-key scores/probabilities [16][128] fp32
constexpr int k4OffTail = k4OffXchg + 2 * 16 * 128 * 4;   // per group: qx[64] | sp[288] | red[8] | part[4][64] floats
constexpr int k4TailFloats = 64 + 288 + 8 + 256;
constexpr int k4OffBar = k4OffTail + 2 * k4TailFloats * 4;
constexpr int k4SmemTotal = k4OffBar + 256 + 1024;

struct Fwd4Params {
    bf16* o;
    float* lse;
    long long o_bs;
    int o_rs;
    int B, H, Tq, Tk, n_main, n_extra, causal, nqb;        // nqb: 128-row query blocks handled here per (b,h)
    float scale_log2e, scale;
    int debug;
    // query rows beyond the last full 128-row block (CLIP: the 257th token) are folded into the same kernel: the
    // group that owns the unit's last tile processes them on the CUDA cores out of the K / V tiles already in
    // shared memory, in the gap where it would otherwise wait for its P.V product
    const bf16* q;
    long long q_bs;
    int q_rs, tail_rows;
};

#ifdef VLK_BRINGUP
__device__ __forceinline__ void dbg4(int enabled, bool who, int g, uint32_t tile, int slot) {
    if (enabled && who && blockIdx.x == 0 && tile >= 2 && tile < 6) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_attn_dbg[(g * 4 + (tile - 2)) * 16 + slot] = t;
    }
}
#else
__device__ __forceinline__ void dbg4(int, bool, int, uint32_t, int) {}
#endif

// swizzled 128-byte row `j` of the K (or V) tile of a stage: keys < 256 in the main tile, the rest in the 16-row box
__device__ __forceinline__ const uint8_t* kv_row(const uint8_t* main_tile, int j) {
    const uint8_t* base = j < 256 ? main_tile : main_tile + 32 * 1024;
    const int r = j & 255;
    return base + (r >> 3) * 1024 + (r & 7) * 128;
}

// One query row against all keys of the unit, by the 128 threads of a softmax group (CUDA cores).
__device__ __forceinline__ void tail_row(const Fwd4Params& p, const uint8_t* st, float* scr, int g, int b, int h, int qi,
                                         int gt /* thread index inside the group */, uint32_t dbg_tile = 0) {
    float* qx = scr;            // [64]   scaled query
    float* sp = scr + 64;       // [288]  scores -> probabilities
    float* red = scr + 352;     // [8]
    float* part = scr + 360;    // [4][64]
    const int lane = gt & 31, w = gt >> 5;
    const int lim = p.causal ? min(p.Tk, qi + (p.Tk - p.Tq) + 1) : p.Tk;
    dbg4(p.debug, gt == 0, g, dbg_tile, 9);
    if (gt < 64) qx[gt] = __bfloat162float(p.q[b * p.q_bs + static_cast<size_t>(qi) * p.q_rs + h * 64 + gt]) * p.scale;
    ptx::named_bar_sync(1 + g, 128);
    dbg4(p.debug, gt == 0, g, dbg_tile, 10);
    float m = -INFINITY;
    for (int j = gt; j < lim; j += 128) {
        const uint8_t* kr = kv_row(st, j);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent chains (one chain of 64 FMAs is 256 clk of latency)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float kf[8];
            unpack8(*reinterpret_cast<const uint4*>(kr + ((c ^ (j & 7)) << 4)), kf);
            const float4 q0 = *reinterpret_cast<const float4*>(qx + c * 8);
            const float4 q1 = *reinterpret_cast<const float4*>(qx + c * 8 + 4);
            a0 = fmaf(q0.x, kf[0], a0); a1 = fmaf(q0.y, kf[1], a1); a2 = fmaf(q0.z, kf[2], a2);
            a3 = fmaf(q0.w, kf[3], a3); a0 = fmaf(q1.x, kf[4], a0); a1 = fmaf(q1.y, kf[5], a1);
            a2 = fmaf(q1.z, kf[6], a2); a3 = fmaf(q1.w, kf[7], a3);
        }
        const float acc = (a0 + a1) + (a2 + a3);
        sp[j] = acc;
        m = fmaxf(m, acc);
    }
    dbg4(p.debug, gt == 0, g, dbg_tile, 11);
    m = warp_max(m);
    if (lane == 0) red[w] = m;
    ptx::named_bar_sync(1 + g, 128);
    m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float sum = 0.f;
    for (int j = gt; j < lim; j += 128) {
        const float e = __expf(sp[j] - m);
        sp[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[4 + w] = sum;
    ptx::named_bar_sync(1 + g, 128);
    sum = red[4] + red[5] + red[6] + red[7];
    dbg4(p.debug, gt == 0, g, dbg_tile, 12);
    // O[d]: warp w walks keys j = w, w+4, ...; lane owns dims 2*lane, 2*lane+1 (conflict-free swizzled reads)
    float a0 = 0.f, a1 = 0.f;
    const uint8_t* vt = st + 34 * 1024;
    {
        // four keys per round with independent loads and accumulators (the serial loop was 65 dependent
        // address -> load -> FMA steps per warp)
        float b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f, d0 = 0.f, d1 = 0.f;
        auto vword = [&](int j) {
            const uint8_t* vr = kv_row(vt, j);
            const uint32_t word = *reinterpret_cast<const uint32_t*>(vr + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4);
            return __bfloat1622float2(*reinterpret_cast<const bf162*>(&word));
        };
        int j = w;
        for (; j + 12 < lim; j += 16) {
            const float2 v0 = vword(j), v1 = vword(j + 4), v2 = vword(j + 8), v3 = vword(j + 12);
            const float p0 = sp[j], p1 = sp[j + 4], p2 = sp[j + 8], p3 = sp[j + 12];
            a0 = fmaf(p0, v0.x, a0); a1 = fmaf(p0, v0.y, a1);
            b0 = fmaf(p1, v1.x, b0); b1 = fmaf(p1, v1.y, b1);
            c0 = fmaf(p2, v2.x, c0); c1 = fmaf(p2, v2.y, c1);
            d0 = fmaf(p3, v3.x, d0); d1 = fmaf(p3, v3.y, d1);
        }
        for (; j < lim; j += 4) {
            const float2 v2 = vword(j);
            const float pj = sp[j];
            a0 = fmaf(pj, v2.x, a0);
            a1 = fmaf(pj, v2.y, a1);
        }
        a0 = (a0 + b0) + (c0 + d0);
        a1 = (a1 + b1) + (c1 + d1);
    }
    dbg4(p.debug, gt == 0, g, dbg_tile, 13);
    part[w * 64 + 2 * lane] = a0;
    part[w * 64 + 2 * lane + 1] = a1;
    ptx::named_bar_sync(1 + g, 128);
    if (gt < 32) {
        const float inv = 1.0f / sum;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            o0 += part[x * 64 + 2 * gt];
            o1 += part[x * 64 + 2 * gt + 1];
        }
        bf16* orow = p.o + b * p.o_bs + static_cast<size_t>(qi) * p.o_rs + h * 64;
        reinterpret_cast<bf162*>(orow)[gt] = __floats2bfloat162_rn(o0 * inv, o1 * inv);
        if (gt == 0 && p.lse != nullptr) p.lse[(static_cast<size_t>(b) * p.H + h) * p.Tq + qi] = m + __logf(sum);
    }
    ptx::named_bar_sync(1 + g, 128);  // scratch may be reused
    dbg4(p.debug, gt == 0, g, dbg_tile, 14);
}

__global__ void __launch_bounds__(384, 1)
attn_fwd_tcgen05_v4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                           const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_kx,
                           const __grid_constant__ CUtensorMap tmap_vx, Fwd4Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k4OffBar);
    uint64_t* kfull = bars;          // [2] per K/V stage
    uint64_t* vfull = bars + 2;      // [2]
    uint64_t* kv_empty = bars + 4;   // [2]  MMA commit + both groups
    uint64_t* q_full = bars + 6;     // [2] per slot
    uint64_t* q_empty = bars + 8;    // [2]
    uint64_t* s_ready = bars + 10;   // [2]
    uint64_t* p_ready = bars + 12;   // [2]
    uint64_t* o_ready = bars + 14;   // [2]
    uint64_t* s_free = bars + 16;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp: uniform for the compiler
    const int n_main = p.n_main, n_extra = p.n_extra, nqb = p.nqb;
    const int num_units = p.B * p.H;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kfull[i], 1);
            ptx::mbar_init(&vfull[i], 1);
            ptx::mbar_init(&kv_empty[i], 3);
            ptx::mbar_init(&q_full[i], 1);
            ptx::mbar_init(&q_empty[i], 1);
            ptx::mbar_init(&s_ready[i], 1);
            ptx::mbar_init(&p_ready[i], 1);
            ptx::mbar_init(&o_ready[i], 1);
            ptx::mbar_init(&s_free[i], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t kv_bytes = static_cast<uint32_t>(n_main) * 128u + (n_extra > 0 ? 2048u : 0u);

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        // converged warp; one elected lane issues (TMA operands stay in uniform registers)
        const bool issuer = ptx::elect_one();
        uint32_t qcnt0 = 0, qcnt1 = 0;
        int it = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
            const int b = u / p.H, h = u % p.H, s = it & 1;
            uint8_t* st = smem + s * k4KVStage;
            ptx::mbar_wait(&kv_empty[s], ((it >> 1) & 1) ^ 1);
            if (issuer) {
                ptx::mbar_arrive_expect_tx(&kfull[s], kv_bytes);
                ptx::tma_load_3d(st, &tmap_k, &kfull[s], h * 64, 0, b);
                if (n_extra > 0) ptx::tma_load_3d(st + 32 * 1024, &tmap_kx, &kfull[s], h * 64, 256, b);
                ptx::mbar_arrive_expect_tx(&vfull[s], kv_bytes);
                ptx::tma_load_3d(st + 34 * 1024, &tmap_v, &vfull[s], h * 64, 0, b);
                if (n_extra > 0) ptx::tma_load_3d(st + 66 * 1024, &tmap_vx, &vfull[s], h * 64, 256, b);
            }
            __syncwarp();
            for (int qb = 0; qb < nqb; ++qb) {
                const int g = qb & 1;
                const uint32_t qc = g ? qcnt1 : qcnt0;
                ptx::mbar_wait(&q_empty[g], (qc & 1) ^ 1);
                if (issuer) {
                    ptx::mbar_arrive_expect_tx(&q_full[g], kQBytes);
                    ptx::tma_load_3d(smem + k4OffQ + g * 16 * 1024, &tmap_q, &q_full[g], h * 64, qb * 128, b);
                }
                __syncwarp();
                if (g) ++qcnt1; else ++qcnt0;
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        // Flat tile sequence t = 0, 1, 2, ... (slot = query block & 1):  issue S(t), then P.V(t-1) as soon as its
        // probabilities are complete — both groups run their softmax concurrently (two warps per scheduler hide each
        // other's TMEM / MUFU latencies; measured against "P.V(t-1) first": 73.5 vs 78.0 us per CLIP layer).
        // The warp walks the loop CONVERGED and one elected lane issues: every tcgen05.mma operand is then warp-uniform
        // and lives in uniform registers (issued from inside `if (lane == 0)` each MMA cost an ELECT / R2UR loop of ~25
        // instructions, 20 MMAs per tile).
        const bool issuer = ptx::elect_one();
        uint32_t cnt_s0 = 0, cnt_s1 = 0, cnt_o0 = 0, cnt_o1 = 0;
        const uint32_t idesc_s = ptx::make_idesc_bf16_f32(128, n_main, 0, 0);
        constexpr uint32_t idesc_o = ptx::make_idesc_bf16_f32(128, 64, 0, 1);
        const int ksteps = n_main / 16;
        int prev_g = -1, prev_s = 0, prev_last = 0;
        uint32_t prev_kvpar = 0;
        auto pv_ready = [&]() {   // warp-uniform: the previous tile's probabilities are complete
            return __all_sync(0xffffffffu, ptx::mbar_try_wait(&p_ready[prev_g], (prev_g ? cnt_o1 : cnt_o0) & 1));
        };
        auto issue_pv = [&]() {
            // P.V of the previous tile (its probabilities are complete)
            const uint32_t sv = ptx::smem_u32(smem + prev_s * k4KVStage) + 34 * 1024;
            ptx::mbar_wait(&vfull[prev_s], prev_kvpar);
            ptx::tc_fence_after_sync();
            const uint32_t t_o = tmem + prev_g * 256 + 128, t_p = tmem + prev_g * 256;
            if (issuer) {
                if (ksteps == 16) {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        ptx::umma_bf16_ts(t_o, t_p + k * 8, ptx::make_smem_desc_sw128(sv + k * 2048, 8192, 1024), idesc_o,
                                          k != 0);
                } else {
                    for (int k = 0; k < ksteps; ++k)
                        ptx::umma_bf16_ts(t_o, t_p + k * 8, ptx::make_smem_desc_sw128(sv + k * 2048, 8192, 1024), idesc_o,
                                          k != 0);
                }
                ptx::umma_commit(&o_ready[prev_g]);
                if (prev_last) ptx::umma_commit(&kv_empty[prev_s]);  // every MMA of that unit has been issued
            }
            __syncwarp();
            if (prev_g) ++cnt_o1; else ++cnt_o0;
        };
        int it = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t kvpar = (it >> 1) & 1;
            const uint32_t sk = ptx::smem_u32(smem + s * k4KVStage);
            ptx::mbar_wait(&kfull[s], kvpar);
            for (int qb = 0; qb < nqb; ++qb) {
                const int g = qb & 1;
                const uint32_t par = (g ? cnt_s1 : cnt_s0) & 1;
                const uint32_t sq = ptx::smem_u32(smem + k4OffQ + g * 16 * 1024);
                // Two things are pending: the scores of this tile (needs its Q and the group's free TMEM slot) and the
                // P.V of the previous tile (needs its probabilities).  Whichever becomes possible first is issued first
                // — with the cheap issue path a P.V no longer delays the next scores noticeably, while waiting for the
                // other group's slot before issuing a finished tile's P.V cost 4-5 us per tile (phase stamps).
                bool s_done = false, pv_done = prev_g < 0;
                while (!s_done || !pv_done) {
                    if (!s_done && __all_sync(0xffffffffu, ptx::mbar_try_wait(&q_full[g], par) &&
                                                               ptx::mbar_try_wait(&s_free[g], par ^ 1))) {
                        ptx::tc_fence_after_sync();
                        if (issuer) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma_bf16_ss(tmem + g * 256, ptx::make_smem_desc_sw128(sq + k * 32, 16, 1024),
                                                  ptx::make_smem_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k != 0);
                            ptx::umma_commit(&s_ready[g]);
                        }
                        __syncwarp();
                        if (g) ++cnt_s1; else ++cnt_s0;
                        s_done = true;
                    }
                    if (!pv_done && pv_ready()) {
                        issue_pv();
                        pv_done = true;
                    }
                }
                prev_g = g;
                prev_s = s;
                prev_kvpar = kvpar;
                prev_last = (qb == nqb - 1);
            }
        }
        if (prev_g >= 0) {  // drain: the last tile's P.V
            ptx::mbar_wait(&p_ready[prev_g], (prev_g ? cnt_o1 : cnt_o0) & 1);
            issue_pv();
        }
    } else if (warp >= 4) {
        // ===================================== softmax / epilogue groups ========================
        const int g = (warp - 4) >> 2;
        const int row = ((warp & 3) << 5) + lane;
        const uint32_t trow = tmem + g * 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        float* spx = reinterpret_cast<float*>(smem + k4OffXchg) + g * 16 * 128;
        const uint8_t* qrow = smem + k4OffQ + g * 16 * 1024 + (row >> 3) * 1024 + (row & 7) * 128;
        const bool leader = (threadIdx.x & 127) == 0;
        uint32_t cnt = 0;
        int it = 0;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
            const int b = u / p.H, h = u % p.H, s = it & 1;
            const uint32_t kvpar = (it >> 1) & 1;
            const uint8_t* st = smem + s * k4KVStage;
            // never run ahead of the unit (a group without a tile in this unit would otherwise arrive on kv_empty
            // for a phase that has not started)
            ptx::mbar_wait(&kfull[s], kvpar);
            for (int qb = g; qb < nqb; qb += 2, ++cnt) {
                const uint32_t par = cnt & 1;
                const int qi = qb * 128 + row;
                int lim = p.Tk;
                if (p.causal) lim = min(p.Tk, qi + (p.Tk - p.Tq) + 1);
                if (lim < 1) lim = 1;
                float m = -INFINITY;
                dbg4(p.debug, leader, g, cnt, 0);
                if (n_extra > 0) {
                    ptx::mbar_wait(&q_full[g], par);
                    ptx::mbar_wait(&kfull[s], kvpar);
#pragma unroll 1
                    for (int e = 0; e < n_extra; ++e) {
                        const float sc = (256 + e < lim) ? dot_q_kx(qrow, row & 7, st + 32 * 1024, e) : -INFINITY;
                        spx[e * 128 + row] = sc;
                        m = fmaxf(m, sc);
                    }
                }
                dbg4(p.debug, leader, g, cnt, 1);
                ptx::mbar_wait(&s_ready[g], par);
                ptx::tc_fence_after_sync();
                dbg4(p.debug, leader, g, cnt, 2);
                const int full = min(__reduce_min_sync(0xffffffffu, lim), n_main) & ~31;
                // software-pipelined TMEM reads: the next 32 columns are in flight while these are reduced
                if (full > 0) {
                    uint32_t ra[32], rb[32];
                    ptx::tmem_ld_32x32b_x32(trow, ra);
                    ptx::tmem_ld_wait();
                    for (int c = 0; c < full; c += 64) {
                        const bool more1 = c + 32 < full, more2 = c + 64 < full;
                        if (more1) ptx::tmem_ld_32x32b_x32(trow + c + 32, rb);
#pragma unroll
                        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(ra[i]));
                        if (more1) {
                            ptx::tmem_ld_wait();
                            if (more2) ptx::tmem_ld_32x32b_x32(trow + c + 64, ra);
#pragma unroll
                            for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(rb[i]));
                            if (more2) ptx::tmem_ld_wait();
                        }
                    }
                }
                for (int c = full; c < n_main; c += 16) {
                    uint32_t r[16];
                    ptx::tmem_ld_32x32b_x16(trow + c, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c + i < lim) m = fmaxf(m, __uint_as_float(r[i]));
                }
                const float mb = m * p.scale_log2e;
                float sum = 0.f;
                dbg4(p.debug, leader, g, cnt, 3);
                if (full > 0) {
                    uint32_t ra[32], rb[32];
                    auto exp_store = [&](const uint32_t (&r)[32], int c) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float e0 = ex2_fast(fmaf(__uint_as_float(r[2 * i]), p.scale_log2e, -mb));
                            const float e1 = ex2_fast(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2e, -mb));
                            sum += e0 + e1;
                            const bf162 h2 = __floats2bfloat162_rn(e0, e1);
                            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                        }
                        // P columns [c/2, c/2+16) overwrite score columns this thread has already consumed
                        ptx::tmem_st_32x32b_x16(trow + (c >> 1), pk);
                    };
                    ptx::tmem_ld_32x32b_x32(trow, ra);
                    ptx::tmem_ld_wait();
                    for (int c = 0; c < full; c += 64) {
                        const bool more1 = c + 32 < full, more2 = c + 64 < full;
                        if (more1) ptx::tmem_ld_32x32b_x32(trow + c + 32, rb);
                        exp_store(ra, c);
                        if (more1) {
                            ptx::tmem_ld_wait();
                            if (more2) ptx::tmem_ld_32x32b_x32(trow + c + 64, ra);
                            exp_store(rb, c + 32);
                            if (more2) ptx::tmem_ld_wait();
                        }
                    }
                }
                for (int c = full; c < n_main; c += 16) {
                    uint32_t r[16];
                    ptx::tmem_ld_32x32b_x16(trow + c, r);
                    ptx::tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float e0 =
                            (c + 2 * i < lim) ? ex2_fast(fmaf(__uint_as_float(r[2 * i]), p.scale_log2e, -mb)) : 0.f;
                        const float e1 = (c + 2 * i + 1 < lim)
                                             ? ex2_fast(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2e, -mb))
                                             : 0.f;
                        sum += e0 + e1;
                        const bf162 h2 = __floats2bfloat162_rn(e0, e1);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                    }
                    ptx::tmem_st_32x32b_x8(trow + (c >> 1), pk);
                }
#pragma unroll 1
                for (int e = 0; e < n_extra; ++e) {
                    const float pe = (256 + e < lim) ? ex2_fast(fmaf(spx[e * 128 + row], p.scale_log2e, -mb)) : 0.f;
                    sum += pe;
                    spx[e * 128 + row] = pe;
                }
                dbg4(p.debug, leader, g, cnt, 4);
                ptx::tmem_st_wait();
                ptx::tc_fence_before_sync();
                ptx::named_bar_sync(1 + g, 128);
                if (leader) {
                    ptx::mbar_arrive(&p_ready[g]);   // P complete -> MMA warp may issue P.V
                    ptx::mbar_arrive(&q_empty[g]);   // this slot's Q tile is no longer needed (S done, dots done)
                }
                dbg4(p.debug, leader, g, cnt, 5);
                // The leftover rows go to group (unit parity), i.e. ALTERNATE between the groups from unit to unit: the phase
                // stamps of the bring-up build (profiles/r02/attn_clip_phase_stamps.md) show them costing 4.5-5 us per
                // unit on the CUDA cores; always on the last tile's group they made that group's chain 10 us against
                // 5 us for the other one, and the MMA issue order couples the two (75.5 -> 65.0 us per CLIP layer).
                if (p.tail_rows > 0 && (it & 1) == g && qb + 2 >= nqb) {
                    // leftover query rows of this (batch, head), while the tensor core works on this tile's P.V
                    ptx::mbar_wait(&vfull[s], kvpar);
                    float* scr = reinterpret_cast<float*>(smem + k4OffTail) + g * k4TailFloats;
                    for (int tr = 0; tr < p.tail_rows; ++tr)
                        tail_row(p, st, scr, g, b, h, nqb * 128 + tr, threadIdx.x & 127, cnt);
                }
                ptx::mbar_wait(&o_ready[g], par);
                if (n_extra > 0) ptx::mbar_wait(&vfull[s], kvpar);
                ptx::tc_fence_after_sync();
                dbg4(p.debug, leader, g, cnt, 6);
                const float inv = 1.0f / sum;
                bf16* orow = p.o + b * p.o_bs + static_cast<size_t>(qi) * p.o_rs + h * 64;
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(trow + 128 + c, r);
                    ptx::tmem_ld_wait();
                    float t[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = __uint_as_float(r[i]);
#pragma unroll 1
                    for (int e = 0; e < n_extra; ++e) {
                        const uint8_t* vrow = st + 66 * 1024 + (e >> 3) * 1024 + (e & 7) * 128;
                        const float pe = spx[e * 128 + row];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float v8[8];
                            unpack8(*reinterpret_cast<const uint4*>(vrow + ((((c >> 3) + q) ^ (e & 7)) << 4)), v8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) t[q * 8 + i] = fmaf(pe, v8[i], t[q * 8 + i]);
                        }
                    }
                    if (qi < p.Tq) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float o8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) o8[i] = t[q * 8 + i] * inv;
                            stg16(orow + c + q * 8, pack8(o8));
                        }
                    }
                }
                if (qi < p.Tq && p.lse != nullptr)
                    p.lse[(static_cast<size_t>(b) * p.H + h) * p.Tq + qi] = m * p.scale + __logf(sum);
                dbg4(p.debug, leader, g, cnt, 7);
                ptx::tc_fence_before_sync();
                ptx::named_bar_sync(1 + g, 128);
                if (leader) ptx::mbar_arrive(&s_free[g]);  // TMEM slot (and spx) may be reused
                dbg4(p.debug, leader, g, cnt, 8);
            }
            // this group is done with the unit's K/V stage (Kx / Vx reads included); also when it had no tile
            ptx::named_bar_sync(1 + g, 128);
            if (leader) ptx::mbar_arrive(&kv_empty[s]);
        }
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// [B, T, W] bf16 view (W contiguous, row stride rs, batch stride bs): box = 64 x box_rows x 1, 128B swizzle.
int make_tmap3(CUtensorMap* map, const void* base, int W, int T, int B, int rs, long long bs, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    VLK_REQUIRE(fn != nullptr, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLK_REQUIRE(r == CUDA_SUCCESS, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
    return VLK_OK;
}

}  // namespace

// 64 < Tk <= 272, Tq >= 64: full 128-row query blocks on the tensor cores; a remainder of <= 8 rows (the 257th CLIP
// token) is folded into the same kernel.  All strides in elements, multiples of 8; bases 16-byte aligned.
int attn_mid_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                 long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs,
                 int causal, float scale, cudaStream_t s) {
    int tail = Tq % 128;
    int tc_rows = Tq - tail;
    if (tail > 8) {  // a longer remainder gets its own (partly empty) 128-row block
        tc_rows = Tq;
        tail = 0;
    }
    static bool configured4 = false;
    if (!configured4) {
        VLK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      k4SmemTotal));
        configured4 = true;
    }
    const int n_main = Tk >= 256 ? 256 : (Tk + 15) / 16 * 16;
    CUtensorMap tq, tk, tv, tkx, tvx;
    int rc = make_tmap3(&tq, q, H * 64, Tq, B, q_rs, q_bs, 128);
    if (rc) return rc;
    rc = make_tmap3(&tk, k, H * 64, Tk, B, k_rs, k_bs, n_main);
    if (rc) return rc;
    rc = make_tmap3(&tv, v, H * 64, Tk, B, v_rs, v_bs, n_main);
    if (rc) return rc;
    rc = make_tmap3(&tkx, k, H * 64, Tk, B, k_rs, k_bs, 16);
    if (rc) return rc;
    rc = make_tmap3(&tvx, v, H * 64, Tk, B, v_rs, v_bs, 16);
    if (rc) return rc;
    Fwd4Params p4;
    p4.o = static_cast<bf16*>(o);
    p4.lse = lse;
    p4.o_bs = o_bs;
    p4.o_rs = o_rs;
    p4.B = B;
    p4.H = H;
    p4.Tq = Tq;
    p4.Tk = Tk;
    p4.n_main = n_main;
    p4.n_extra = Tk > 256 ? Tk - 256 : 0;
    p4.causal = causal;
    p4.nqb = (tc_rows + 127) / 128;
    p4.scale = scale;
    p4.scale_log2e = scale * 1.4426950408889634f;
#ifdef VLK_BRINGUP
    p4.debug = getenv("VLK_ATTN_DEBUG") != nullptr;
#else
    p4.debug = 0;
#endif
    p4.q = static_cast<const bf16*>(q);
    p4.q_bs = q_bs;
    p4.q_rs = q_rs;
    p4.tail_rows = tail;   // folded into the persistent kernel
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_attn_fwd: no sm_100 device");
    const int units = B * H;
    attn_fwd_tcgen05_v4_kernel<<<units < sms ? units : sms, 384, k4SmemTotal, s>>>(tq, tk, tv, tkx, tvx, p4);
    VLK_CHECK_LAUNCH("vlk_attn_fwd(tcgen05 mid)");
    return VLK_OK;
}

}  // namespace vlk

#ifdef VLK_BRINGUP
// bring-up only (not part of include/vlk.h, not in the shipped library): copy the attention phase stamps to the host
extern "C" int vlk_debug_dump(long long* host_out, int n) {
    if (n > 64 * 16) n = 64 * 16;
    cudaError_t e = cudaMemcpyFromSymbol(host_out, vlk::g_attn_dbg, sizeof(long long) * n);
    return static_cast<int>(e);
}
#endif
