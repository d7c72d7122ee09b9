// Scaled-dot-product attention FORWARD, head dim 64, on the CUDA cores (fp32 math, online softmax) — the shapes the
// tensor-core kernels do not take: a few query rows against a long key range (KV-cached decode beyond 64 cached
// tokens; one warp group per row) and the general streaming kernel for Tq < 64 with Tk > 64.
// Everything else runs on tcgen05: attention_pair.cu (Tq, Tk <= 64, forward + backward), attention_tcgen05.cu
// (64 < Tk <= 272, the CLIP tower), attention_flash.cu (longer sequences, forward + backward).
//
// Addressing: element (b, t, h, d) of Q lives at q + b*q_bs + t*q_rs + h*64 + d (same for K, V, O), so packed
// c_attn / kv_proj / in_proj outputs are consumed without a head transpose.
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
}  // namespace vlk
