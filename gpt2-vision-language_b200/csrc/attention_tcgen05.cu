// Attention forward on the tcgen05 tensor cores for mid-length sequences (64 < Tk <= 272: the CLIP ViT-L/14
// encoder has 257 tokens, 16 heads x 64).  One CTA = 128 query rows of one (batch, head):
//
//   TMA:  Q [128 x 64], K [Tk x 64], V [Tk x 64]  -> shared memory (128B swizzle, zero-filled past the sequence)
//   MMA1: S = Q K^T            (M=128, N=Tk padded to 16, K=64; fp32 in TMEM, 272 columns)
//   softmax: thread t owns TMEM lane t = one query row -> no cross-thread reduction at all;
//            row max, exp2, row sum in registers; un-normalised P written as bf16 into a swizzled K-major
//            shared-memory tile
//   MMA2: O = P V              (M=128, N=64, K=Tk; V is fed as an MN-major B operand, i.e. no transpose)
//   epilogue: O / rowsum -> bf16 -> global; optional log-sum-exp for the backward pass.
//
// The whole key range fits one tile, so there is no online-softmax rescaling here; sequences longer than 272
// use the streaming kernel in attention_simt.cu.
#include <cuda.h>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace vlk {

int attn_fwd_simt(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                  long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                  int o_rs, int causal, float scale, cudaStream_t stream, int q_row0);

bool attn_pair_applicable(int B, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs,
                          int v_rs, long long o_bs, int o_rs);
int attn_pair_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                  long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                  int o_rs, int causal, float scale, float dropout_p, const unsigned long long* seed_state,
                  unsigned int stream_id, cudaStream_t stream);
bool attn_small_applicable(int Tq, int Tk);
int attn_small_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                   long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                   int o_rs, int causal, float scale, float dropout_p, const unsigned long long* seed_state,
                   unsigned int stream_id, cudaStream_t stream);

int attn_flash_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                   long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                   int o_rs, int causal, float scale, cudaStream_t stream);

// bring-up instrumentation: per-phase %globaltimer stamps of the first CTAs (read back by vlk_debug_dump)
__device__ long long g_attn_dbg[64 * 16];
__device__ __forceinline__ void dbg_stamp(int enabled, int slot) {
    if (enabled && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.z < 32 && blockIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_attn_dbg[blockIdx.z * 16 + slot] = t;
    }
}

namespace {

constexpr int kMaxKeys = 272;
constexpr int kQBytes = 128 * 128;                 // 128 rows x 64 bf16
constexpr int kKVBytes = 35 * 1024;                // >= 272 rows x 128 B, multiple of 1024
constexpr int kPSlabBytes = 128 * 128;             // 128 rows x 64 keys
constexpr int kPSlabs = (kMaxKeys + 63) / 64;      // 5
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kKVBytes;
constexpr int kOffP = kOffV + kKVBytes;
constexpr int kOffBar = kOffP + kPSlabs * kPSlabBytes;
constexpr int kSmemBytes = kOffBar + 64 + 1024;    // barriers + tmem slot + alignment slack
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemO = 320;                   // O accumulator columns [320, 384)

struct FwdParams {
    bf16* o;
    float* lse;
    long long o_bs;
    int o_rs;
    int H, Tq, Tk, tkp, box_rows, causal;
    float scale_log2e, scale;
};

__global__ void __launch_bounds__(128, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* bar_v = bar_qk + 1;
    uint64_t* bar_s = bar_qk + 2;
    uint64_t* bar_o = bar_qk + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int tkp = p.tkp;
    const int n1 = tkp > 256 ? 256 : tkp, n2 = tkp - n1;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        ptx::mbar_init(bar_qk, 1);
        ptx::mbar_init(bar_v, 1);
        ptx::mbar_init(bar_s, 1);
        ptx::mbar_init(bar_o, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (threadIdx.x == 0) {
        const int nbox = tkp / p.box_rows;  // 1 or 2
        const uint32_t kv_bytes = static_cast<uint32_t>(tkp) * 128u;
        ptx::mbar_arrive_expect_tx(bar_qk, kQBytes + kv_bytes);
        ptx::tma_load_3d(smem + kOffQ, &tmap_q, bar_qk, h * 64, q0, b);
        for (int i = 0; i < nbox; ++i)
            ptx::tma_load_3d(smem + kOffK + i * p.box_rows * 128, &tmap_k, bar_qk, h * 64, i * p.box_rows, b);
        ptx::mbar_arrive_expect_tx(bar_v, kv_bytes);
        for (int i = 0; i < nbox; ++i)
            ptx::tma_load_3d(smem + kOffV + i * p.box_rows * 128, &tmap_v, bar_v, h * 64, i * p.box_rows, b);

        // ---- S = Q K^T ----
        ptx::mbar_wait(bar_qk, 0);
        ptx::tc_fence_after_sync();
        const uint32_t sq = ptx::smem_u32(smem + kOffQ), sk = ptx::smem_u32(smem + kOffK);
        const uint32_t idesc1 = ptx::make_idesc_bf16_f32(128, n1, 0, 0);
        const uint32_t idesc2 = ptx::make_idesc_bf16_f32(128, n2 > 0 ? n2 : 16, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sq + k * 32, 16, 1024);
            ptx::umma_bf16_ss(tmem, da, ptx::make_smem_desc_sw128(sk + k * 32, 16, 1024), idesc1, k != 0);
            if (n2 > 0)
                ptx::umma_bf16_ss(tmem + 256, da, ptx::make_smem_desc_sw128(sk + 256 * 128 + k * 32, 16, 1024), idesc2,
                                  k != 0);
        }
        ptx::umma_commit(bar_s);
    }

    // ---- softmax: thread = query row = TMEM lane ----
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after_sync();
    const int row = threadIdx.x;                    // local query row
    const int qi = q0 + row;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    int lim = p.Tk;                                 // keys [0, lim) are visible to this row
    if (p.causal) lim = min(p.Tk, qi + (p.Tk - p.Tq) + 1);
    if (lim < 1) lim = 1;                           // rows past Tq: keep the math finite, result is discarded
    float m = -INFINITY;
    for (int c = 0; c < tkp; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(trow + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c + i < lim) m = fmaxf(m, __uint_as_float(r[i]));
    }
    const float mb = m * p.scale_log2e;
    float sum = 0.f;
    uint8_t* prow = smem + kOffP + (row >> 3) * 1024 + (row & 7) * 128;
    for (int c = 0; c < tkp; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(trow + c, r);
        ptx::tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            e[i] = (c + i < lim) ? exp2f(__uint_as_float(r[i]) * p.scale_log2e - mb) : 0.f;
        }
        // round to bf16 first so that the row sum matches what the tensor core will multiply
        uint4 lo, hi;
        {
            float t0[8], t1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                t0[i] = e[i];
                t1[i] = e[8 + i];
            }
            lo = pack8(t0);
            hi = pack8(t1);
            float u0[8], u1[8];
            unpack8(lo, u0);
            unpack8(hi, u1);
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += u0[i] + u1[i];
        }
        const int slab = c >> 6, chunk = (c & 63) >> 3;  // 16-byte chunk index inside the 128-byte row
        uint8_t* base = prow + slab * kPSlabBytes;
        *reinterpret_cast<uint4*>(base + ((chunk ^ (row & 7)) << 4)) = lo;
        *reinterpret_cast<uint4*>(base + (((chunk + 1) ^ (row & 7)) << 4)) = hi;
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy), then CTA-wide hand-off
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before_sync();
    __syncthreads();

    if (threadIdx.x == 0) {
        // ---- O = P V ----
        ptx::mbar_wait(bar_v, 0);
        ptx::tc_fence_after_sync();
        const uint32_t sp = ptx::smem_u32(smem + kOffP), sv = ptx::smem_u32(smem + kOffV);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);  // B (= V) is MN-major
        const int ksteps = tkp / 16;
        for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sp + (k >> 2) * kPSlabBytes + (k & 3) * 32, 16, 1024);
            const uint64_t db = ptx::make_smem_desc_sw128(sv + k * 2048, 8192, 1024);
            ptx::umma_bf16_ss(tmem + kTmemO, da, db, idesc, k != 0);
        }
        ptx::umma_commit(bar_o);
    }
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after_sync();
    const float inv = 1.0f / sum;
    if (qi < p.Tq) {
        bf16* orow = p.o + b * p.o_bs + static_cast<size_t>(qi) * p.o_rs + h * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(trow + kTmemO + c, r);
            ptx::tmem_ld_wait();
            float t0[8], t1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                t0[i] = __uint_as_float(r[i]) * inv;
                t1[i] = __uint_as_float(r[8 + i]) * inv;
            }
            stg16(orow + c, pack8(t0));
            stg16(orow + c + 8, pack8(t1));
        }
        if (p.lse != nullptr) p.lse[(static_cast<size_t>(b) * p.H + h) * p.Tq + qi] = m * p.scale + __logf(sum);
    } else {
        // keep the warp converged for the .sync.aligned TMEM loads of its other lanes
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(trow + kTmemO + c, r);
            ptx::tmem_ld_wait();
        }
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// v2: probabilities stay in TENSOR MEMORY.  After the row max, each thread overwrites its own S row in place
// with bf16 P (two values per 32-bit column, tcgen05.st) and O = P V runs with A read from TMEM.  That removes the
// 80 KB shared-memory P tile: a CTA needs 84 KB of smem and 256 TMEM columns, so TWO CTAs fit per SM and one
// CTA's softmax overlaps the other's loads and MMAs.  Keys beyond 256 (CLIP has 257) are not worth a second
// 256-column accumulator: their scores and their P.V contribution (<= 16 keys) are computed on the CUDA cores.
//   TMEM columns: S fp32 [0,256) -> P bf16x2 [0,128) in place; O fp32 [128,192) (dead S columns).
// ------------------------------------------------------------------------------------------------
constexpr int k2OffQ = 0;                  // 16 KB
constexpr int k2OffK = 16 * 1024;          // 32 KB  (256 keys)
constexpr int k2OffKx = 48 * 1024;         //  2 KB  (keys 256..271)
constexpr int k2OffV = 50 * 1024;          // 32 KB
constexpr int k2OffVx = 82 * 1024;         //  2 KB
constexpr int k2OffBar = 84 * 1024;
constexpr int k2SmemBytes = k2OffBar + 64 + 1024;
constexpr uint32_t k2TmemCols = 256;
constexpr uint32_t k2TmemO = 128;

struct Fwd2Params {
    bf16* o;
    float* lse;
    long long o_bs;
    int o_rs;
    int H, Tq, Tk, n_main, n_extra, causal;
    float scale_log2e, scale;
    int debug;
};

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

// 256 threads: two threads per query row (warps w and w+4 share TMEM lane quadrant w&3).  Thread "half" h owns
// score columns [128h, 128h+128): it reads them, writes its P columns in place inside its OWN column range
// (keys 0..127 -> TMEM cols [0,64), keys 128..255 -> cols [128,192)) and later normalises O columns [32h, 32h+32)
// (O lives in cols [192,256), dead score columns of half 1).  Row max / row sum / extra-key probabilities are
// exchanged through shared memory.
constexpr int k2OffXchg = k2OffBar + 128;                 // smax[2][128] | ssum[2][128] | spx[16][128] floats
constexpr int k2XchgBytes = (2 + 2 + 16) * 128 * 4;
constexpr int k2SmemTotal = k2OffXchg + k2XchgBytes + 1024;
constexpr uint32_t k2TmemO3 = 192;

__global__ void __launch_bounds__(256, 2)
attn_fwd_tcgen05_v2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                           const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_kx,
                           const __grid_constant__ CUtensorMap tmap_vx, Fwd2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + k2OffBar);
    uint64_t* bar_v = bar_qk + 1;
    uint64_t* bar_s = bar_qk + 2;
    uint64_t* bar_o = bar_qk + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);
    float* smax = reinterpret_cast<float*>(smem + k2OffXchg);
    float* ssum = smax + 256;
    float* spx = ssum + 256;

    const int warp = threadIdx.x >> 5;
    const int half = threadIdx.x >> 7;
    const int row = threadIdx.x & 127;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int n_main = p.n_main, n_extra = p.n_extra;
    dbg_stamp(p.debug, 0);

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        ptx::mbar_init(bar_qk, 1);
        ptx::mbar_init(bar_v, 1);
        ptx::mbar_init(bar_s, 1);
        ptx::mbar_init(bar_o, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, k2TmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    dbg_stamp(p.debug, 1);

    if (threadIdx.x == 0) {
        const uint32_t kv_bytes = static_cast<uint32_t>(n_main) * 128u + (n_extra > 0 ? 2048u : 0u);
        ptx::mbar_arrive_expect_tx(bar_qk, kQBytes + kv_bytes);
        ptx::tma_load_3d(smem + k2OffQ, &tmap_q, bar_qk, h * 64, q0, b);
        ptx::tma_load_3d(smem + k2OffK, &tmap_k, bar_qk, h * 64, 0, b);
        if (n_extra > 0) ptx::tma_load_3d(smem + k2OffKx, &tmap_kx, bar_qk, h * 64, 256, b);
        ptx::mbar_arrive_expect_tx(bar_v, kv_bytes);
        ptx::tma_load_3d(smem + k2OffV, &tmap_v, bar_v, h * 64, 0, b);
        if (n_extra > 0) ptx::tma_load_3d(smem + k2OffVx, &tmap_vx, bar_v, h * 64, 256, b);

        ptx::mbar_wait(bar_qk, 0);
        dbg_stamp(p.debug, 2);
        ptx::tc_fence_after_sync();
        const uint32_t sq = ptx::smem_u32(smem + k2OffQ), sk = ptx::smem_u32(smem + k2OffK);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, n_main, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(tmem, ptx::make_smem_desc_sw128(sq + k * 32, 16, 1024),
                              ptx::make_smem_desc_sw128(sk + k * 32, 16, 1024), idesc, k != 0);
        ptx::umma_commit(bar_s);
    }

    const int qi = q0 + row;
    const uint32_t trow = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    int lim = p.Tk;
    if (p.causal) lim = min(p.Tk, qi + (p.Tk - p.Tq) + 1);
    if (lim < 1) lim = 1;
    const uint8_t* qrow = smem + k2OffQ + (row >> 3) * 1024 + (row & 7) * 128;
    // this thread's score columns [dom0, dom1); `full`: end of the prefix that needs no masking.  Warp-uniform on
    // purpose: the tcgen05.ld/st below are .sync.aligned, every lane must run the same trip counts.
    const int dom0 = min(half * 128, n_main), dom1 = min(dom0 + 128, n_main);
    const int lim_min = __reduce_min_sync(0xffffffffu, lim);
    const int full = dom0 + (max(min(lim_min, dom1) - dom0, 0) & ~31);
    const uint32_t pcol0 = half * 128;  // P columns of this half start here (in place, inside its own range)

    // scores of the extra keys (CUDA cores; half 0 only), needed for the row max.  They are parked in this
    // thread's private slots of the exchange buffer (a rolled loop: 16 unrolled copies of the dot product blew
    // the instruction cache).
    float m = -INFINITY;
    if (half == 0 && n_extra > 0) {
        ptx::mbar_wait(bar_qk, 0);  // Q / Kx are in shared memory
#pragma unroll 1
        for (int e = 0; e < n_extra; ++e) {
            const float sc = (256 + e < lim) ? dot_q_kx(qrow, row & 7, smem + k2OffKx, e) : -INFINITY;
            spx[e * 128 + row] = sc;
            m = fmaxf(m, sc);
        }
    }
    dbg_stamp(p.debug, 3);
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after_sync();
    dbg_stamp(p.debug, 4);
    for (int c = dom0; c < full; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(trow + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
    }
    for (int c = full; c < dom1; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(trow + c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c + i < lim) m = fmaxf(m, __uint_as_float(r[i]));
    }
    smax[half * 128 + row] = m;
    __syncthreads();
    m = fmaxf(smax[row], smax[128 + row]);
    const float mb = m * p.scale_log2e;
    float sum = 0.f;
    dbg_stamp(p.debug, 5);
    for (int c = dom0; c < full; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(trow + c, r);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float e0 = ex2_fast(fmaf(__uint_as_float(r[2 * i]), p.scale_log2e, -mb));
            const float e1 = ex2_fast(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2e, -mb));
            sum += e0 + e1;
            const bf162 h2 = __floats2bfloat162_rn(e0, e1);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        // P columns overwrite score columns of this thread's own range that it has already consumed
        ptx::tmem_st_32x32b_x16(trow + pcol0 + ((c - dom0) >> 1), pk);
    }
    for (int c = full; c < dom1; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld_32x32b_x16(trow + c, r);
        ptx::tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float e0 = (c + 2 * i < lim) ? ex2_fast(fmaf(__uint_as_float(r[2 * i]), p.scale_log2e, -mb)) : 0.f;
            const float e1 =
                (c + 2 * i + 1 < lim) ? ex2_fast(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2e, -mb)) : 0.f;
            sum += e0 + e1;
            const bf162 h2 = __floats2bfloat162_rn(e0, e1);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        ptx::tmem_st_32x32b_x8(trow + pcol0 + ((c - dom0) >> 1), pk);
    }
    if (half == 0) {
#pragma unroll 1
        for (int e = 0; e < n_extra; ++e) {
            const float pe = (256 + e < lim) ? ex2_fast(fmaf(spx[e * 128 + row], p.scale_log2e, -mb)) : 0.f;
            sum += pe;
            spx[e * 128 + row] = pe;
        }
    }
    ssum[half * 128 + row] = sum;
    dbg_stamp(p.debug, 6);
    ptx::tmem_st_wait();
    ptx::tc_fence_before_sync();
    __syncthreads();
    dbg_stamp(p.debug, 7);

    if (threadIdx.x == 0) {
        ptx::mbar_wait(bar_v, 0);
        ptx::tc_fence_after_sync();
        const uint32_t sv = ptx::smem_u32(smem + k2OffV);
        const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);  // A = P (TMEM, K-major), B = V (MN-major)
        const int ksteps = n_main / 16;
        for (int k = 0; k < ksteps; ++k) {
            const uint32_t pa = k < 8 ? k * 8 : 128 + (k - 8) * 8;       // where that key block's P lives
            ptx::umma_bf16_ts(tmem + k2TmemO3, tmem + pa, ptx::make_smem_desc_sw128(sv + k * 2048, 8192, 1024), idesc,
                              k != 0);
        }
        ptx::umma_commit(bar_o);
    }
    ptx::mbar_wait(bar_v, 0);  // Vx visible to every thread
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after_sync();
    dbg_stamp(p.debug, 8);
    sum = ssum[row] + ssum[128 + row];
    const float inv = 1.0f / sum;
    {
        // this thread normalises and stores O columns [32*half, 32*half + 32) of its row
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(trow + k2TmemO3 + half * 32, r);
        ptx::tmem_ld_wait();
        float t[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = __uint_as_float(r[i]);
        for (int e = 0; e < n_extra; ++e) {  // rank-1 updates from the extra keys
            const uint8_t* vrow = smem + k2OffVx + (e >> 3) * 1024 + (e & 7) * 128;
            const float pe = spx[e * 128 + row];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v8[8];
                unpack8(*reinterpret_cast<const uint4*>(vrow + (((half * 4 + q) ^ (e & 7)) << 4)), v8);
#pragma unroll
                for (int i = 0; i < 8; ++i) t[q * 8 + i] = fmaf(pe, v8[i], t[q * 8 + i]);
            }
        }
        if (qi < p.Tq) {
            bf16* orow = p.o + b * p.o_bs + static_cast<size_t>(qi) * p.o_rs + h * 64 + half * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o8[i] = t[q * 8 + i] * inv;
                stg16(orow + q * 8, pack8(o8));
            }
        }
    }
    if (half == 0 && qi < p.Tq && p.lse != nullptr)
        p.lse[(static_cast<size_t>(b) * p.H + h) * p.Tq + qi] = m * p.scale + __logf(sum);
    dbg_stamp(p.debug, 9);
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, k2TmemCols);
    }
    dbg_stamp(p.debug, 10);
}

// ------------------------------------------------------------------------------------------------
// v4: persistent, warp-specialised pipeline (one CTA per SM).  Work unit = one (batch, head): its K / V tiles are
// loaded ONCE (double-buffered across units) and shared by all of its 128-row query blocks; two query blocks are in
// flight at a time, each in its own 256-column TMEM slot with its own 128-thread softmax group, so one block's
// exp / store phase overlaps the other's MMAs and epilogue and the next unit's TMA loads.
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7: group 0   warps 8-11: group 1
// Per-slot TMEM layout as in v2: S fp32 [0,256) -> P bf16x2 in place [0,128); O fp32 [128,192).
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
    // MMA issue order: 0 = wait for the previous tile's probabilities before producing the next tile's scores (the two
    // groups' exp phases never overlap); 1 = scores first, so both groups run their softmax concurrently (two warps
    // per scheduler hide each other's TMEM / MUFU latencies)
    int order;
};

__device__ __forceinline__ void dbg4(int enabled, bool who, int g, uint32_t tile, int slot) {
    if (enabled && who && blockIdx.x == 0 && tile >= 2 && tile < 6) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_attn_dbg[(g * 4 + (tile - 2)) * 16 + slot] = t;
    }
}

// swizzled 128-byte row `j` of the K (or V) tile of a stage: keys < 256 in the main tile, the rest in the 16-row box
__device__ __forceinline__ const uint8_t* kv_row(const uint8_t* main_tile, int j) {
    const uint8_t* base = j < 256 ? main_tile : main_tile + 32 * 1024;
    const int r = j & 255;
    return base + (r >> 3) * 1024 + (r & 7) * 128;
}

// One query row against all keys of the unit, by the 128 threads of a softmax group (CUDA cores).
__device__ __forceinline__ void tail_row(const Fwd4Params& p, const uint8_t* st, float* scr, int g, int b, int h, int qi,
                                         int gt /* thread index inside the group */) {
    float* qx = scr;            // [64]   scaled query
    float* sp = scr + 64;       // [288]  scores -> probabilities
    float* red = scr + 352;     // [8]
    float* part = scr + 360;    // [4][64]
    const int lane = gt & 31, w = gt >> 5;
    const int lim = p.causal ? min(p.Tk, qi + (p.Tk - p.Tq) + 1) : p.Tk;
    if (gt < 64) qx[gt] = __bfloat162float(p.q[b * p.q_bs + static_cast<size_t>(qi) * p.q_rs + h * 64 + gt]) * p.scale;
    ptx::named_bar_sync(1 + g, 128);
    float m = -INFINITY;
    for (int j = gt; j < lim; j += 128) {
        const uint8_t* kr = kv_row(st, j);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float kf[8];
            unpack8(*reinterpret_cast<const uint4*>(kr + ((c ^ (j & 7)) << 4)), kf);
            const float4 q0 = *reinterpret_cast<const float4*>(qx + c * 8);
            const float4 q1 = *reinterpret_cast<const float4*>(qx + c * 8 + 4);
            acc = fmaf(q0.x, kf[0], acc); acc = fmaf(q0.y, kf[1], acc); acc = fmaf(q0.z, kf[2], acc);
            acc = fmaf(q0.w, kf[3], acc); acc = fmaf(q1.x, kf[4], acc); acc = fmaf(q1.y, kf[5], acc);
            acc = fmaf(q1.z, kf[6], acc); acc = fmaf(q1.w, kf[7], acc);
        }
        sp[j] = acc;
        m = fmaxf(m, acc);
    }
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
    // O[d]: warp w walks keys j = w, w+4, ...; lane owns dims 2*lane, 2*lane+1 (conflict-free swizzled reads)
    float a0 = 0.f, a1 = 0.f;
    const uint8_t* vt = st + 34 * 1024;
    for (int j = w; j < lim; j += 4) {
        const uint8_t* vr = kv_row(vt, j);
        const uint32_t word = *reinterpret_cast<const uint32_t*>(vr + (((lane >> 2) ^ (j & 7)) << 4) + (lane & 3) * 4);
        const float2 v2 = __bfloat1622float2(*reinterpret_cast<const bf162*>(&word));
        const float pj = sp[j];
        a0 = fmaf(pj, v2.x, a0);
        a1 = fmaf(pj, v2.y, a1);
    }
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        if (lane == 0) {
            uint32_t qcnt[2] = {0, 0};
            int it = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
                const int b = u / p.H, h = u % p.H, s = it & 1;
                uint8_t* st = smem + s * k4KVStage;
                ptx::mbar_wait(&kv_empty[s], ((it >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&kfull[s], kv_bytes);
                ptx::tma_load_3d(st, &tmap_k, &kfull[s], h * 64, 0, b);
                if (n_extra > 0) ptx::tma_load_3d(st + 32 * 1024, &tmap_kx, &kfull[s], h * 64, 256, b);
                ptx::mbar_arrive_expect_tx(&vfull[s], kv_bytes);
                ptx::tma_load_3d(st + 34 * 1024, &tmap_v, &vfull[s], h * 64, 0, b);
                if (n_extra > 0) ptx::tma_load_3d(st + 66 * 1024, &tmap_vx, &vfull[s], h * 64, 256, b);
                for (int qb = 0; qb < nqb; ++qb) {
                    const int g = qb & 1;
                    ptx::mbar_wait(&q_empty[g], (qcnt[g] & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&q_full[g], kQBytes);
                    ptx::tma_load_3d(smem + k4OffQ + g * 16 * 1024, &tmap_q, &q_full[g], h * 64, qb * 128, b);
                    ++qcnt[g];
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        // Ping-pong schedule over the flat tile sequence t = 0, 1, 2, ... (slot = query block & 1):
        //     wait softmax(t-1) done  ->  issue S(t)  ->  issue P.V(t-1)
        // so the exp-heavy (MUFU-bound) phase of tile t never overlaps that of tile t-1; instead it runs under
        // the P.V product, normalisation and stores of tile t-1 in the other group.
        if (lane == 0) {
            uint32_t cnt_s[2] = {0, 0}, cnt_o[2] = {0, 0};
            const uint32_t idesc_s = ptx::make_idesc_bf16_f32(128, n_main, 0, 0);
            const uint32_t idesc_o = ptx::make_idesc_bf16_f32(128, 64, 0, 1);
            const int ksteps = n_main / 16;
            int prev_g = -1, prev_s = 0, prev_last = 0;
            uint32_t prev_kvpar = 0;
            auto issue_pv = [&]() {
                // P.V of the previous tile (its softmax was already waited for)
                const uint32_t sv = ptx::smem_u32(smem + prev_s * k4KVStage) + 34 * 1024;
                ptx::mbar_wait(&vfull[prev_s], prev_kvpar);
                ptx::tc_fence_after_sync();
                for (int k = 0; k < ksteps; ++k)
                    ptx::umma_bf16_ts(tmem + prev_g * 256 + 128, tmem + prev_g * 256 + k * 8,
                                      ptx::make_smem_desc_sw128(sv + k * 2048, 8192, 1024), idesc_o, k != 0);
                ptx::umma_commit(&o_ready[prev_g]);
                ++cnt_o[prev_g];
                if (prev_last) ptx::umma_commit(&kv_empty[prev_s]);  // every MMA of that unit has been issued
            };
            int it = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
                const int s = it & 1;
                const uint32_t kvpar = (it >> 1) & 1;
                const uint32_t sk = ptx::smem_u32(smem + s * k4KVStage);
                ptx::mbar_wait(&kfull[s], kvpar);
                for (int qb = 0; qb < nqb; ++qb) {
                    const int g = qb & 1;
                    // only when the previous tile's probabilities are complete may the next tile's scores be produced
                    // (keeps the two groups' exp phases from overlapping); S(t) goes first because it is on the
                    // critical path, P.V(t-1) right behind it
                    const uint32_t par = cnt_s[g] & 1;
                    if (p.order != 0) {
                        auto issue_s = [&]() {
                            ptx::tc_fence_after_sync();
                            const uint32_t sq1 = ptx::smem_u32(smem + k4OffQ + g * 16 * 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma_bf16_ss(tmem + g * 256, ptx::make_smem_desc_sw128(sq1 + k * 32, 16, 1024),
                                                  ptx::make_smem_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k != 0);
                            ptx::umma_commit(&s_ready[g]);
                            ++cnt_s[g];
                        };
                        if (p.order == 1) {          // scores first, then the previous tile's P.V
                            ptx::mbar_wait(&q_full[g], par);
                            ptx::mbar_wait(&s_free[g], par ^ 1);
                            issue_s();
                            if (prev_g >= 0) {
                                ptx::mbar_wait(&p_ready[prev_g], cnt_o[prev_g] & 1);
                                issue_pv();
                            }
                        } else {                     // whichever becomes ready first
                            bool s_done = false, pv_done = prev_g < 0;
                            while (!s_done || !pv_done) {
                                if (!s_done && ptx::mbar_try_wait(&q_full[g], par) &&
                                    ptx::mbar_try_wait(&s_free[g], par ^ 1)) {
                                    issue_s();
                                    s_done = true;
                                }
                                if (!pv_done && ptx::mbar_try_wait(&p_ready[prev_g], cnt_o[prev_g] & 1)) {
                                    issue_pv();
                                    pv_done = true;
                                }
                            }
                        }
                        prev_g = g;
                        prev_s = s;
                        prev_kvpar = kvpar;
                        prev_last = (qb == nqb - 1);
                        continue;
                    }
                    bool pv_pending = prev_g >= 0;
                    if (pv_pending) {
                        ptx::mbar_wait(&p_ready[prev_g], cnt_o[prev_g] & 1);
                        // if this tile's inputs are not there yet, do not let the finished tile's P.V queue behind them
                        if (!(ptx::mbar_try_wait(&q_full[g], par) && ptx::mbar_try_wait(&s_free[g], par ^ 1))) {
                            issue_pv();
                            pv_pending = false;
                        }
                    }
                    ptx::mbar_wait(&q_full[g], par);
                    ptx::mbar_wait(&s_free[g], par ^ 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t sq = ptx::smem_u32(smem + k4OffQ + g * 16 * 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16_ss(tmem + g * 256, ptx::make_smem_desc_sw128(sq + k * 32, 16, 1024),
                                          ptx::make_smem_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k != 0);
                    ptx::umma_commit(&s_ready[g]);
                    ++cnt_s[g];
                    if (pv_pending) issue_pv();
                    prev_g = g;
                    prev_s = s;
                    prev_kvpar = kvpar;
                    prev_last = (qb == nqb - 1);
                }
            }
            if (prev_g >= 0) {  // drain: the last tile's P.V
                ptx::mbar_wait(&p_ready[prev_g], cnt_o[prev_g] & 1);
                issue_pv();
            }
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
                if (p.tail_rows > 0 && qb == nqb - 1) {
                    // leftover query rows of this (batch, head), while the tensor core works on this tile's P.V
                    ptx::mbar_wait(&vfull[s], kvpar);
                    float* scr = reinterpret_cast<float*>(smem + k4OffTail) + g * k4TailFloats;
                    for (int tr = 0; tr < p.tail_rows; ++tr)
                        tail_row(p, st, scr, g, b, h, nqb * 128 + tr, threadIdx.x & 127);
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
}  // namespace vlk

using namespace vlk;

extern "C" int vlk_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq,
                            int Tk, long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs,
                            long long o_bs, int o_rs, int causal, float scale, float dropout_p,
                            const unsigned long long* seed_state, unsigned int stream_id, void* stream) {
    VLK_REQUIRE(q && k && v && o, VLK_ERR_INVALID_ARG, "vlk_attn_fwd: null pointer");
    VLK_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && (dropout_p == 0.f || seed_state), VLK_ERR_INVALID_ARG,
                "vlk_attn_fwd: dropout_p=%f needs a seed state", dropout_p);
    VLK_REQUIRE(dropout_p == 0.f || attn_small_applicable(Tq, Tk), VLK_ERR_UNSUPPORTED,
                "vlk_attn_fwd: attention dropout is only implemented for Tq, Tk <= 64 (the Q-Former shapes)");
    VLK_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, VLK_ERR_INVALID_ARG, "vlk_attn_fwd: B=%d H=%d Tq=%d Tk=%d", B, H,
                Tq, Tk);
    VLK_REQUIRE(q_rs % 8 == 0 && k_rs % 8 == 0 && v_rs % 8 == 0 && o_rs % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 &&
                    v_bs % 8 == 0 && o_bs % 8 == 0,
                VLK_ERR_ALIGNMENT, "vlk_attn_fwd: strides must be multiples of 8 elements");
    VLK_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), VLK_ERR_ALIGNMENT,
                "vlk_attn_fwd: 16B alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const char* force = getenv("VLK_ATTN_IMPL");
    if (attn_small_applicable(Tq, Tk) &&
        (dropout_p > 0.f || !(force && (strcmp(force, "simt") == 0 || strcmp(force, "flash") == 0)))) {
        // two heads per CTA on the tensor cores; VLK_ATTN_IMPL=small keeps the CUDA-core kernel for cross-checks
        if (!(force && strcmp(force, "small") == 0) &&
            attn_pair_applicable(B, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs))
            return attn_pair_fwd(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal,
                                 scale, dropout_p, seed_state, stream_id, s);
        return attn_small_fwd(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal,
                              scale, dropout_p, seed_state, stream_id, s);
    }
    // long sequences (GPT-2 pretraining, T = 1024): streaming tcgen05 kernel
    if ((Tk > kMaxKeys && !(force && strcmp(force, "simt") == 0)) || (force && strcmp(force, "flash") == 0))
        return attn_flash_fwd(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal,
                              scale, s);
    const bool want_tc = force ? (strncmp(force, "tcgen05", 7) == 0) : (Tk > 64);
    if (!want_tc || Tk > kMaxKeys || Tq < 64 || (force && strcmp(force, "simt") == 0))
        return attn_fwd_simt(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal,
                             scale, s, 0);

    // full 128-row query blocks on the tensor cores; a short tail (e.g. the 257th CLIP token) on the CUDA cores
    int tail = Tq % 128;
    int tc_rows = Tq - tail;
    if (tail > 8) {  // a longer remainder gets its own (partly empty) 128-row block
        tc_rows = Tq;
        tail = 0;
    }
    if (!(force && strcmp(force, "tcgen05v1") == 0)) {
        static bool configured2 = false;
        if (!configured2) {
            VLK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          k2SmemTotal));
            configured2 = true;
        }
        Fwd2Params p2;
        p2.o = static_cast<bf16*>(o);
        p2.lse = lse;
        p2.o_bs = o_bs;
        p2.o_rs = o_rs;
        p2.H = H;
        p2.Tq = Tq;
        p2.Tk = Tk;
        p2.n_main = Tk >= 256 ? 256 : (Tk + 15) / 16 * 16;
        p2.n_extra = Tk > 256 ? Tk - 256 : 0;
        p2.causal = causal;
        p2.scale = scale;
        p2.scale_log2e = scale * 1.4426950408889634f;
        p2.debug = getenv("VLK_ATTN_DEBUG") != nullptr ? atoi(getenv("VLK_ATTN_DEBUG")) : 0;
        CUtensorMap tq, tk, tv, tkx, tvx;
        int rc = make_tmap3(&tq, q, H * 64, Tq, B, q_rs, q_bs, 128);
        if (rc) return rc;
        rc = make_tmap3(&tk, k, H * 64, Tk, B, k_rs, k_bs, p2.n_main);
        if (rc) return rc;
        rc = make_tmap3(&tv, v, H * 64, Tk, B, v_rs, v_bs, p2.n_main);
        if (rc) return rc;
        rc = make_tmap3(&tkx, k, H * 64, Tk, B, k_rs, k_bs, 16);
        if (rc) return rc;
        rc = make_tmap3(&tvx, v, H * 64, Tk, B, v_rs, v_bs, 16);
        if (rc) return rc;
        if (!(force && strcmp(force, "tcgen05v3") == 0)) {
            // persistent pipeline: one CTA per SM walks the (batch, head) units
            static bool configured4 = false;
            if (!configured4) {
                VLK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              k4SmemTotal));
                configured4 = true;
            }
            Fwd4Params p4;
            p4.o = p2.o;
            p4.lse = p2.lse;
            p4.o_bs = o_bs;
            p4.o_rs = o_rs;
            p4.B = B;
            p4.H = H;
            p4.Tq = Tq;
            p4.Tk = Tk;
            p4.n_main = p2.n_main;
            p4.n_extra = p2.n_extra;
            p4.causal = causal;
            p4.nqb = (tc_rows + 127) / 128;
            p4.scale = scale;
            p4.scale_log2e = p2.scale_log2e;
            p4.debug = p2.debug;
            p4.q = static_cast<const bf16*>(q);
            p4.q_bs = q_bs;
            p4.q_rs = q_rs;
            p4.tail_rows = tail;   // folded into the persistent kernel
            // measured at B=64, H=16, T=257 (graph-timed): order 0 78.0 us, 1 73.5 us, 2 76.5 us
            p4.order = 1;
            if (const char* f = getenv("VLK_ATTN_V4_ORDER")) p4.order = atoi(f);
            tail = 0;
            const int sms = device_sm_count();
            VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_attn_fwd: no sm_100 device");
            const int units = B * H;
            attn_fwd_tcgen05_v4_kernel<<<units < sms ? units : sms, 384, k4SmemTotal, s>>>(tq, tk, tv, tkx, tvx, p4);
            VLK_CHECK_LAUNCH("vlk_attn_fwd(tcgen05 v4)");
        } else {
            const dim3 grid2((tc_rows + 127) / 128, H, B);
            attn_fwd_tcgen05_v2_kernel<<<grid2, 256, k2SmemTotal, s>>>(tq, tk, tv, tkx, tvx, p2);
            VLK_CHECK_LAUNCH("vlk_attn_fwd(tcgen05 v3)");
        }
        if (tail > 0)
            return attn_fwd_simt(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal,
                                 scale, s, tc_rows);
        return VLK_OK;
    }
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured = true;
    }
    FwdParams p;
    p.o = static_cast<bf16*>(o);
    p.lse = lse;
    p.o_bs = o_bs;
    p.o_rs = o_rs;
    p.H = H;
    p.Tq = Tq;
    p.Tk = Tk;
    p.tkp = (Tk + 15) / 16 * 16;
    p.box_rows = p.tkp > 256 ? p.tkp / 2 : p.tkp;
    p.causal = causal;
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    CUtensorMap tq, tk, tv;
    int rc = make_tmap3(&tq, q, H * 64, Tq, B, q_rs, q_bs, 128);
    if (rc) return rc;
    rc = make_tmap3(&tk, k, H * 64, Tk, B, k_rs, k_bs, p.box_rows);
    if (rc) return rc;
    rc = make_tmap3(&tv, v, H * 64, Tk, B, v_rs, v_bs, p.box_rows);
    if (rc) return rc;
    const dim3 grid((tc_rows + 127) / 128, H, B);
    attn_fwd_tcgen05_kernel<<<grid, 128, kSmemBytes, s>>>(tq, tk, tv, p);
    VLK_CHECK_LAUNCH("vlk_attn_fwd(tcgen05)");
    if (tail > 0)
        return attn_fwd_simt(q, k, v, o, lse, B, H, Tq, Tk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs, causal,
                             scale, s, tc_rows);
    return VLK_OK;
}

// bring-up only (not part of include/vlk.h): copy the attention phase stamps to the host
extern "C" int vlk_debug_dump(long long* host_out, int n) {
    if (n > 64 * 16) n = 64 * 16;
    cudaError_t e = cudaMemcpyFromSymbol(host_out, vlk::g_attn_dbg, sizeof(long long) * n);
    return static_cast<int>(e);
}
