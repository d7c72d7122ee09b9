// Scaled-dot-product attention, head dim 64, on the CUDA cores (fp32 math, online softmax).
// This is the general kernel family: any Tq/Tk, causal or not, forward AND backward.  It serves the short
// sequences of the captioning step (T = 31..64, launch-bound, a few GFLOP in total) and every backward pass.
// The long forward case (CLIP, 257 keys) is served by the tcgen05 kernel in attention_tcgen05.cu.
//
// Addressing: element (b, t, h, d) of Q lives at q + b*q_bs + t*q_rs + h*64 + d (same for K, V, O and the
// gradients), so packed c_attn / kv_proj / in_proj outputs are consumed without a head transpose.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vlk {
namespace {

constexpr int D = 64;          // head dim (fixed on this path: 768/12 = 1024/16 = 64)
constexpr int TILE = 64;       // rows of the "other side" staged in shared memory per iteration
constexpr int LDS = 66;        // padded row stride (bf16 elements) -> conflict-free column walks
constexpr int kWarps = 4;
constexpr int ROWS_PER_BLOCK = 16;  // rows of the "own side" handled by one block (4 per warp)

struct Addr {
    long long bs;
    int rs;
};

// Stage `rows` x 64 bf16 from global (row stride rs) into padded shared memory; rows >= valid are zeroed.
__device__ __forceinline__ void stage_tile(bf16* __restrict__ dst, const bf16* __restrict__ src, int rs, int valid) {
    for (int v = threadIdx.x; v < TILE * 8; v += blockDim.x) {
        const int r = v >> 3, c = (v & 7) * 8;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (r < valid) u = ldg16(src + static_cast<size_t>(r) * rs + c);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + r * LDS + c);
        d32[0] = u.x;
        d32[1] = u.y;
        d32[2] = u.z;
        d32[3] = u.w;
    }
}

// dot of a broadcast fp32 vector (shared, 64 floats) with one padded bf16 row (shared)
__device__ __forceinline__ float dot64(const float* __restrict__ qv, const bf16* __restrict__ row) {
    float acc = 0.f;
    const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
    for (int w = 0; w < D / 2; ++w) {
        const float2 q2 = *reinterpret_cast<const float2*>(qv + 2 * w);
        const uint32_t kw = r32[w];
        const float2 k2 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&kw));
        acc = fmaf(q2.x, k2.x, acc);
        acc = fmaf(q2.y, k2.y, acc);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
attn_fwd_simt_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                     bf16* __restrict__ o, float* __restrict__ lse, int H, int Tq, int Tk, Addr qa, Addr ka, Addr va,
                     Addr oa, int causal, float scale, int q_row0) {
    __shared__ __align__(16) bf16 sK[TILE * LDS];
    __shared__ __align__(16) bf16 sV[TILE * LDS];
    __shared__ __align__(16) float sQ[ROWS_PER_BLOCK][D];
    __shared__ float sP[kWarps][TILE];

    const int b = blockIdx.z, h = blockIdx.y;
    const int q0 = q_row0 + blockIdx.x * ROWS_PER_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bf16* qb = q + b * qa.bs + h * D;
    const bf16* kb = k + b * ka.bs + h * D;
    const bf16* vb = v + b * va.bs + h * D;
    const int shift = Tk - Tq;  // causal: query i sees keys j <= i + shift

    for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * D; idx += blockDim.x) {
        const int r = idx / D, d = idx % D;
        sQ[r][d] = (q0 + r < Tq) ? __bfloat162float(qb[static_cast<size_t>(q0 + r) * qa.rs + d]) * scale : 0.f;
    }

    constexpr int RPW = ROWS_PER_BLOCK / kWarps;
    float m[RPW], l[RPW], acc0[RPW], acc1[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        m[r] = -INFINITY;
        l[r] = 0.f;
        acc0[r] = 0.f;
        acc1[r] = 0.f;
    }
    int k_end = Tk;
    if (causal) k_end = min(Tk, q0 + ROWS_PER_BLOCK + shift);
    for (int kt = 0; kt < k_end; kt += TILE) {
        __syncthreads();
        stage_tile(sK, kb + static_cast<size_t>(kt) * ka.rs, ka.rs, min(TILE, Tk - kt));
        stage_tile(sV, vb + static_cast<size_t>(kt) * va.rs, va.rs, min(TILE, Tk - kt));
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int qi = q0 + warp * RPW + r;
            if (qi >= Tq) continue;  // warp-uniform
            const int lim = causal ? min(Tk, qi + shift + 1) : Tk;  // keys < lim are visible
            float s0 = (kt + lane < lim) ? dot64(sQ[warp * RPW + r], sK + lane * LDS) : -INFINITY;
            float s1 = (kt + 32 + lane < lim) ? dot64(sQ[warp * RPW + r], sK + (32 + lane) * LDS) : -INFINITY;
            const float mt = warp_max(fmaxf(s0, s1));
            if (mt == -INFINITY) continue;  // whole tile masked for this row
            const float mn = fmaxf(m[r], mt);
            const float corr = __expf(m[r] - mn);
            const float p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);
            l[r] = l[r] * corr + warp_sum(p0 + p1);
            m[r] = mn;
            __syncwarp();
            sP[warp][lane] = p0;
            sP[warp][32 + lane] = p1;
            __syncwarp();
            float a0 = acc0[r] * corr, a1 = acc1[r] * corr;
            const int jmax = min(TILE, lim - kt);
            for (int j = 0; j < jmax; ++j) {
                const float p = sP[warp][j];
                const uint32_t vw = reinterpret_cast<const uint32_t*>(sV + j * LDS)[lane];
                const float2 v2 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&vw));
                a0 = fmaf(p, v2.x, a0);
                a1 = fmaf(p, v2.y, a1);
            }
            acc0[r] = a0;
            acc1[r] = a1;
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int qi = q0 + warp * RPW + r;
        if (qi >= Tq) continue;
        const float inv = l[r] > 0.f ? 1.0f / l[r] : 0.f;
        bf16* orow = o + b * oa.bs + static_cast<size_t>(qi) * oa.rs + h * D;
        reinterpret_cast<bf162*>(orow)[lane] = __floats2bfloat162_rn(acc0[r] * inv, acc1[r] * inv);
        if (lse != nullptr && lane == 0) lse[(static_cast<size_t>(b) * H + h) * Tq + qi] = m[r] + __logf(l[r]);
    }
}

// ---------------------------------------------------------------------------------------------------
// backward, query side: delta_i = dO_i . O_i ;  dQ_i = scale * sum_j dS_ij K_j
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_dq_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                   const bf16* __restrict__ o, const bf16* __restrict__ d_o, const float* __restrict__ lse,
                   bf16* __restrict__ dq, float* __restrict__ delta, int H, int Tq, int Tk, Addr qa, Addr ka,
                   Addr va, Addr oa, Addr dqa, int causal, float scale) {
    __shared__ __align__(16) bf16 sK[TILE * LDS];
    __shared__ __align__(16) bf16 sV[TILE * LDS];
    __shared__ __align__(16) float sQ[ROWS_PER_BLOCK][D];
    __shared__ __align__(16) float sdO[ROWS_PER_BLOCK][D];
    __shared__ float sP[kWarps][TILE];

    const int b = blockIdx.z, h = blockIdx.y;
    const int q0 = blockIdx.x * ROWS_PER_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bf16* qb = q + b * qa.bs + h * D;
    const bf16* kb = k + b * ka.bs + h * D;
    const bf16* vb = v + b * va.bs + h * D;
    const bf16* ob = o + b * oa.bs + h * D;
    const bf16* dob = d_o + b * oa.bs + h * D;
    const int shift = Tk - Tq;

    for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * D; idx += blockDim.x) {
        const int r = idx / D, d = idx % D;
        const bool ok = q0 + r < Tq;
        sQ[r][d] = ok ? __bfloat162float(qb[static_cast<size_t>(q0 + r) * qa.rs + d]) * scale : 0.f;
        sdO[r][d] = ok ? __bfloat162float(dob[static_cast<size_t>(q0 + r) * oa.rs + d]) : 0.f;
    }
    __syncthreads();
    constexpr int RPW = ROWS_PER_BLOCK / kWarps;
    float dl[RPW], ls[RPW], acc0[RPW], acc1[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int lr = warp * RPW + r, qi = q0 + lr;
        acc0[r] = acc1[r] = 0.f;
        dl[r] = 0.f;
        ls[r] = 0.f;
        if (qi < Tq) {
            const float2 o2 = __bfloat1622float2(reinterpret_cast<const bf162*>(ob + static_cast<size_t>(qi) * oa.rs)[lane]);
            dl[r] = warp_sum(o2.x * sdO[lr][2 * lane] + o2.y * sdO[lr][2 * lane + 1]);
            ls[r] = lse[(static_cast<size_t>(b) * H + h) * Tq + qi];
            if (lane == 0) delta[(static_cast<size_t>(b) * H + h) * Tq + qi] = dl[r];
        }
    }
    int k_end = Tk;
    if (causal) k_end = min(Tk, q0 + ROWS_PER_BLOCK + shift);
    for (int kt = 0; kt < k_end; kt += TILE) {
        __syncthreads();
        stage_tile(sK, kb + static_cast<size_t>(kt) * ka.rs, ka.rs, min(TILE, Tk - kt));
        stage_tile(sV, vb + static_cast<size_t>(kt) * va.rs, va.rs, min(TILE, Tk - kt));
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int lr = warp * RPW + r, qi = q0 + lr;
            if (qi >= Tq) continue;
            const int lim = causal ? min(Tk, qi + shift + 1) : Tk;
            if (kt >= lim) continue;
            float ds[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int j = hh * 32 + lane;
                ds[hh] = 0.f;
                if (kt + j < lim) {
                    const float s = dot64(sQ[lr], sK + j * LDS);
                    const float p = __expf(s - ls[r]);
                    const float dp = dot64(sdO[lr], sV + j * LDS);
                    ds[hh] = p * (dp - dl[r]);
                }
            }
            __syncwarp();
            sP[warp][lane] = ds[0];
            sP[warp][32 + lane] = ds[1];
            __syncwarp();
            const int jmax = min(TILE, lim - kt);
            float a0 = acc0[r], a1 = acc1[r];
            for (int j = 0; j < jmax; ++j) {
                const float w = sP[warp][j];
                const uint32_t kw = reinterpret_cast<const uint32_t*>(sK + j * LDS)[lane];
                const float2 k2 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&kw));
                a0 = fmaf(w, k2.x, a0);
                a1 = fmaf(w, k2.y, a1);
            }
            acc0[r] = a0;
            acc1[r] = a1;
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int qi = q0 + warp * RPW + r;
        if (qi >= Tq) continue;
        bf16* row = dq + b * dqa.bs + static_cast<size_t>(qi) * dqa.rs + h * D;
        reinterpret_cast<bf162*>(row)[lane] = __floats2bfloat162_rn(acc0[r] * scale, acc1[r] * scale);
    }
}

// ---------------------------------------------------------------------------------------------------
// backward, key side: dV_j = sum_i P_ij dO_i ;  dK_j = scale * sum_i dS_ij Q_i
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_dkv_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                    const bf16* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                    bf16* __restrict__ dk, bf16* __restrict__ dv, int H, int Tq, int Tk, Addr qa, Addr ka, Addr va,
                    Addr oa, Addr dka, Addr dva, int causal, float scale) {
    __shared__ __align__(16) bf16 sQ[TILE * LDS];
    __shared__ __align__(16) bf16 sdO[TILE * LDS];
    __shared__ __align__(16) float sK[ROWS_PER_BLOCK][D];
    __shared__ __align__(16) float sV[ROWS_PER_BLOCK][D];
    __shared__ float sLse[TILE], sDelta[TILE];
    __shared__ float sP[kWarps][TILE], sDS[kWarps][TILE];

    const int b = blockIdx.z, h = blockIdx.y;
    const int k0 = blockIdx.x * ROWS_PER_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bf16* qb = q + b * qa.bs + h * D;
    const bf16* kb = k + b * ka.bs + h * D;
    const bf16* vb = v + b * va.bs + h * D;
    const bf16* dob = d_o + b * oa.bs + h * D;
    const int shift = Tk - Tq;

    for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * D; idx += blockDim.x) {
        const int r = idx / D, d = idx % D;
        const bool ok = k0 + r < Tk;
        sK[r][d] = ok ? __bfloat162float(kb[static_cast<size_t>(k0 + r) * ka.rs + d]) * scale : 0.f;
        sV[r][d] = ok ? __bfloat162float(vb[static_cast<size_t>(k0 + r) * va.rs + d]) : 0.f;
    }
    constexpr int RPW = ROWS_PER_BLOCK / kWarps;
    float dk0[RPW], dk1[RPW], dv0[RPW], dv1[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) dk0[r] = dk1[r] = dv0[r] = dv1[r] = 0.f;

    // causal: key j is seen by queries i >= j - shift
    int q_begin = 0;
    if (causal) q_begin = max(0, k0 - shift) / TILE * TILE;
    for (int qt = q_begin; qt < Tq; qt += TILE) {
        __syncthreads();
        stage_tile(sQ, qb + static_cast<size_t>(qt) * qa.rs, qa.rs, min(TILE, Tq - qt));
        stage_tile(sdO, dob + static_cast<size_t>(qt) * oa.rs, oa.rs, min(TILE, Tq - qt));
        for (int i = threadIdx.x; i < TILE; i += blockDim.x) {
            const bool ok = qt + i < Tq;
            sLse[i] = ok ? lse[(static_cast<size_t>(b) * H + h) * Tq + qt + i] : 0.f;
            sDelta[i] = ok ? delta[(static_cast<size_t>(b) * H + h) * Tq + qt + i] : 0.f;
        }
        __syncthreads();
        const int imax = min(TILE, Tq - qt);
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int lr = warp * RPW + r, kj = k0 + lr;
            if (kj >= Tk) continue;
            const int first = causal ? max(0, kj - shift) : 0;  // first query index that sees key kj
            if (qt + imax <= first) continue;
            float pv[2], dsv[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int i = hh * 32 + lane;
                pv[hh] = 0.f;
                dsv[hh] = 0.f;
                if (i < imax && qt + i >= first) {
                    const float s = dot64(sK[lr], sQ + i * LDS);  // scale folded into sK
                    const float p = __expf(s - sLse[i]);
                    const float dp = dot64(sV[lr], sdO + i * LDS);
                    pv[hh] = p;
                    dsv[hh] = p * (dp - sDelta[i]);
                }
            }
            __syncwarp();
            sP[warp][lane] = pv[0];
            sP[warp][32 + lane] = pv[1];
            sDS[warp][lane] = dsv[0];
            sDS[warp][32 + lane] = dsv[1];
            __syncwarp();
            float a0 = dk0[r], a1 = dk1[r], c0 = dv0[r], c1 = dv1[r];
            for (int i = max(0, first - qt); i < imax; ++i) {
                const float p = sP[warp][i], w = sDS[warp][i];
                const uint32_t qw = reinterpret_cast<const uint32_t*>(sQ + i * LDS)[lane];
                const uint32_t gw = reinterpret_cast<const uint32_t*>(sdO + i * LDS)[lane];
                const float2 q2 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&qw));
                const float2 g2 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&gw));
                a0 = fmaf(w, q2.x, a0);
                a1 = fmaf(w, q2.y, a1);
                c0 = fmaf(p, g2.x, c0);
                c1 = fmaf(p, g2.y, c1);
            }
            dk0[r] = a0;
            dk1[r] = a1;
            dv0[r] = c0;
            dv1[r] = c1;
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int kj = k0 + warp * RPW + r;
        if (kj >= Tk) continue;
        bf16* rk = dk + b * dka.bs + static_cast<size_t>(kj) * dka.rs + h * D;
        bf16* rv = dv + b * dva.bs + static_cast<size_t>(kj) * dva.rs + h * D;
        reinterpret_cast<bf162*>(rk)[lane] = __floats2bfloat162_rn(dk0[r] * scale, dk1[r] * scale);
        reinterpret_cast<bf162*>(rv)[lane] = __floats2bfloat162_rn(dv0[r], dv1[r]);
    }
}

// ---------------------------------------------------------------------------------------------------
// forward for a handful of query rows (the 257th CLIP token left over after the 128-row tensor-core blocks):
// one WARP per (batch, head, row).  Lanes stride the keys for the scores (each K row is one 128-byte line),
// softmax by warp shuffles, then each lane owns two output dims and walks the keys with coalesced V reads.
// ---------------------------------------------------------------------------------------------------
constexpr int kFewMaxKeys = 512;

// One CTA (4 warps) per (batch, head, row): the keys are split over the warps (so the serial chain per warp is
// short), each warp runs an independent softmax over its slice — lane = key for the scores, lane = two output dims
// for P.V with the V rows of 8 keys fetched per round — and the four partial (max, sum, acc) triples are merged
// through shared memory.
__global__ void __launch_bounds__(128)
attn_fwd_fewrows_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                        bf16* __restrict__ o, float* __restrict__ lse, int B, int H, int Tq, int Tk, Addr qa, Addr ka,
                        Addr va, Addr oa, int causal, float scale, int q_row0) {
    __shared__ float sm_m[4], sm_l[4], sm_acc[4][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nrows = Tq - q_row0;
    const int unit = blockIdx.x;
    const int qi = q_row0 + unit % nrows;
    const int h = (unit / nrows) % H, b = unit / (nrows * H);
    const bf16* qrow = q + b * qa.bs + static_cast<size_t>(qi) * qa.rs + h * D;
    const bf16* kb = k + b * ka.bs + h * D;
    const bf16* vb = v + b * va.bs + h * D;
    const int lim = causal ? min(Tk, qi + (Tk - Tq) + 1) : Tk;
    // this warp's key slice: whole groups of 32 keys
    const int groups = (lim + 31) / 32, gpw = (groups + 3) / 4;
    const int g0 = w * gpw, g1 = min(groups, g0 + gpw);
    float qf[D];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float t[8];
        unpack8(ldg16(qrow + c * 8), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) qf[c * 8 + i] = t[i] * scale;
    }
    constexpr int kMaxG = kFewMaxKeys / 32 / 4;  // key groups per warp
    float s[kMaxG];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxG; ++j) {
        const int key = (g0 + j) * 32 + lane;
        s[j] = -INFINITY;
        if (g0 + j < g1 && key < lim) {
            const bf16* kr = kb + static_cast<size_t>(key) * ka.rs;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float t[8];
                unpack8(ldg16(kr + c * 8), t);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(qf[c * 8 + i], t[i], acc);
            }
            s[j] = acc;
            m = fmaxf(m, acc);
        }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxG; ++j) {
        s[j] = (s[j] > -INFINITY) ? __expf(s[j] - m) : 0.f;
        sum += s[j];
    }
    sum = warp_sum(sum);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxG; ++j) {
        if (g0 + j >= g1) break;
        const int base = (g0 + j) * 32;
        const int n = min(32, lim - base);
#pragma unroll
        for (int t0 = 0; t0 < 32; t0 += 8) {
            if (t0 >= n) break;
            float2 vv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {   // eight independent row fetches in flight
                vv[u] = make_float2(0.f, 0.f);
                if (t0 + u < n)
                    vv[u] = __bfloat1622float2(
                        reinterpret_cast<const bf162*>(vb + static_cast<size_t>(base + t0 + u) * va.rs)[lane]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float pj = __shfl_sync(0xffffffffu, s[j], t0 + u);
                a0 = fmaf(pj, vv[u].x, a0);
                a1 = fmaf(pj, vv[u].y, a1);
            }
        }
    }
    if (lane == 0) {
        sm_m[w] = m;
        sm_l[w] = sum;
    }
    sm_acc[w][2 * lane] = a0;
    sm_acc[w][2 * lane + 1] = a1;
    __syncthreads();
    if (w == 0) {
        float M = fmaxf(fmaxf(sm_m[0], sm_m[1]), fmaxf(sm_m[2], sm_m[3]));
        float L = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const float f = sm_m[x] > -INFINITY ? __expf(sm_m[x] - M) : 0.f;
            L += sm_l[x] * f;
            o0 += sm_acc[x][2 * lane] * f;
            o1 += sm_acc[x][2 * lane + 1] * f;
        }
        const float inv = L > 0.f ? 1.0f / L : 0.f;
        bf16* orow = o + b * oa.bs + static_cast<size_t>(qi) * oa.rs + h * D;
        reinterpret_cast<bf162*>(orow)[lane] = __floats2bfloat162_rn(o0 * inv, o1 * inv);
        if (lse != nullptr && lane == 0) lse[(static_cast<size_t>(b) * H + h) * Tq + qi] = M + __logf(L);
    }
}

}  // namespace

int attn_fwd_simt(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                  long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                  int o_rs, int causal, float scale, cudaStream_t stream, int q_row0) {
    // rows [q_row0, Tq) of every (batch, head)
    if (Tq - q_row0 <= 8 && Tk <= kFewMaxKeys) {
        attn_fwd_fewrows_kernel<<<B * H * (Tq - q_row0), 128, 0, stream>>>(
            static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
            static_cast<bf16*>(o), lse, B, H, Tq, Tk, Addr{q_bs, q_rs}, Addr{k_bs, k_rs}, Addr{v_bs, v_rs},
            Addr{o_bs, o_rs}, causal, scale, q_row0);
        VLK_CHECK_LAUNCH("vlk_attn_fwd(fewrows)");
        return VLK_OK;
    }
    const dim3 grid((Tq - q_row0 + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, H, B);
    attn_fwd_simt_kernel<<<grid, kWarps * 32, 0, stream>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v), static_cast<bf16*>(o),
        lse, H, Tq, Tk, Addr{q_bs, q_rs}, Addr{k_bs, k_rs}, Addr{v_bs, v_rs}, Addr{o_bs, o_rs}, causal, scale, q_row0);
    VLK_CHECK_LAUNCH("vlk_attn_fwd(simt)");
    return VLK_OK;
}

bool attn_pair_applicable(int B, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs,
                          int v_rs, long long o_bs, int o_rs);
int attn_pair_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                  void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                  int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                  long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float dropout_p,
                  const unsigned long long* seed_state, unsigned int stream_id, cudaStream_t stream);
bool attn_small_applicable(int Tq, int Tk);
int attn_small_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                   int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                   long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float dropout_p,
                   const unsigned long long* seed_state, unsigned int stream_id, cudaStream_t stream);

int attn_flash_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                   int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                   long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float* delta,
                   cudaStream_t stream);

}  // namespace vlk

using namespace vlk;

extern "C" int vlk_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                            const float* lse, void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk,
                            long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs,
                            long long o_bs, int o_rs, long long dq_bs, int dq_rs, long long dk_bs, int dk_rs,
                            long long dv_bs, int dv_rs, int causal, float scale, float* delta, float dropout_p,
                            const unsigned long long* seed_state, unsigned int stream_id, void* stream) {
    VLK_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || seed_state), VLK_ERR_INVALID_ARG,
                "vlk_attn_bwd: dropout_p=%f needs a seed state", dropout_p);
    VLK_REQUIRE(dropout_p == 0.f || attn_small_applicable(Tq, Tk), VLK_ERR_UNSUPPORTED,
                "vlk_attn_bwd: attention dropout is only implemented for Tq, Tk <= 64 (the Q-Former shapes)");
    VLK_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv && delta, VLK_ERR_INVALID_ARG,
                "vlk_attn_bwd: null pointer");
    VLK_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, VLK_ERR_INVALID_ARG, "vlk_attn_bwd: B=%d H=%d Tq=%d Tk=%d", B, H,
                Tq, Tk);
    VLK_REQUIRE(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && o_rs % 8 == 0 && dq_rs % 2 == 0 &&
                    dk_rs % 2 == 0 && dv_rs % 2 == 0,
                VLK_ERR_ALIGNMENT, "vlk_attn_bwd: row strides must be multiples of 8 elements");
    VLK_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(d_o), VLK_ERR_ALIGNMENT,
                "vlk_attn_bwd: 16B alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        const char* force = getenv("VLK_ATTN_IMPL");
        if (attn_small_applicable(Tq, Tk) &&
            (dropout_p > 0.f || !(force && (strcmp(force, "simt") == 0 || strcmp(force, "flash") == 0)))) {
            if (!(force && strcmp(force, "small") == 0) &&
                attn_pair_applicable(B, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs) && dq_rs % 8 == 0 &&
                dk_rs % 8 == 0 && dv_rs % 8 == 0 && dq_bs % 8 == 0 && dk_bs % 8 == 0 && dv_bs % 8 == 0 && aligned16(dq) &&
                aligned16(dk) && aligned16(dv))
                return attn_pair_bwd(q, k, v, o, d_o, lse, dq, dk, dv, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs,
                                     o_bs, o_rs, dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, causal, scale, dropout_p,
                                     seed_state, stream_id, s);
            return attn_small_bwd(q, k, v, o, d_o, lse, dq, dk, dv, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs,
                                  o_bs, o_rs, dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, causal, scale, dropout_p,
                                  seed_state, stream_id, s);
        }
    }
    {
        // everything longer than the one-CTA-per-head tile runs on the tensor cores (streaming tcgen05 backward)
        const char* force = getenv("VLK_ATTN_IMPL");
        const bool aligned = q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && o_rs % 8 == 0 && dq_rs % 8 == 0 &&
                             dk_rs % 8 == 0 && dv_rs % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 &&
                             o_bs % 8 == 0;
        if (aligned && !(force && strcmp(force, "simt") == 0))
            return attn_flash_bwd(q, k, v, o, d_o, lse, dq, dk, dv, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs,
                                  o_rs, dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs, causal, scale, delta, s);
    }
    float* scratch = delta;
    const dim3 gq((Tq + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, H, B);
    attn_bwd_dq_kernel<<<gq, kWarps * 32, 0, s>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
        static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), lse, static_cast<bf16*>(dq), scratch, H, Tq, Tk,
        Addr{q_bs, q_rs}, Addr{k_bs, k_rs}, Addr{v_bs, v_rs}, Addr{o_bs, o_rs}, Addr{dq_bs, dq_rs}, causal, scale);
    VLK_CHECK_LAUNCH("vlk_attn_bwd(dq)");
    const dim3 gk((Tk + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK, H, B);
    attn_bwd_dkv_kernel<<<gk, kWarps * 32, 0, s>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
        static_cast<const bf16*>(d_o), lse, scratch, static_cast<bf16*>(dk), static_cast<bf16*>(dv), H, Tq, Tk,
        Addr{q_bs, q_rs}, Addr{k_bs, k_rs}, Addr{v_bs, v_rs}, Addr{o_bs, o_rs}, Addr{dk_bs, dk_rs},
        Addr{dv_bs, dv_rs}, causal, scale);
    VLK_CHECK_LAUNCH("vlk_attn_bwd(dkv)");
    return VLK_OK;
}
