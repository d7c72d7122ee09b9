// Streaming (flash) attention on the tcgen05 tensor cores for long sequences (GPT-2 pretraining, T = 1024),
// forward and backward, head dim 64, causal or not.  Three kernels, all with 128 threads = 128 TMEM lanes:
//
//   forward   CTA = 128 queries of one (b,h); loop over 128-key blocks (double-buffered TMA loads):
//             S = Q K_j^T -> TMEM; thread = query row: running max / sum, P (bf16) written in place over S;
//             O_j = P V_j (A from TMEM, V as MN-major B) -> TMEM; o_reg = o_reg * corr + O_j in registers.
//   bwd dK/dV CTA = 128 keys; loop over query blocks.  Everything is computed TRANSPOSED so that the operand that
//             must come from TMEM is always the A operand:  S^T = K_j Q_i^T and dP^T = V_j dO_i^T (thread = key row,
//             lse / delta indexed by column), P^T and dS^T (bf16, in place) are then the A operands of
//             dV_j += P^T dO_i and dK_j += dS^T Q_i.  One smem tile [query x 64] serves both as the K-major B
//             operand of the first two products and as the MN-major B operand of the last two.
//   bwd dQ    CTA = 128 queries; loop over key blocks: S = Q_i K_j^T, dP = dO_i V_j^T, dS in place,
//             dQ_i += dS K_j (K_j as MN-major B).  Recomputing S and dP here (7 instead of 5 block products)
//             buys a backward without atomics.
//
// Scores are kept unscaled in TMEM; probabilities are exp2(s * scale*log2e - lse*log2e) with the forward's lse.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace vlk {
namespace {

constexpr int BQ = 128;   // query rows per block = TMEM lanes
constexpr int BK = 128;   // keys per block
constexpr int kTile = 128 * 128;  // bytes of a 128-row x 64-col bf16 tile

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifdef VLK_BRINGUP
// bring-up instrumentation (never compiled into the shipped library): cycles per phase of the forward loop, summed over
// the iterations of a CTA, by thread 0 and thread 255; read back by vlk_debug_flash_dump
__device__ long long g_flash_dbg[64 * 16];
#define FDBG_DECL long long fd_t = clock64(), fd_t0 = fd_t, fd_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define FDBG(slot) do { const long long t__ = clock64(); fd_acc[slot] += t__ - fd_t; fd_t = t__; } while (0)
#define FDBG_DUMP(iters) do { if (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) < 32 && (threadIdx.x == 0 || threadIdx.x == 255)) { \
    long long* o__ = g_flash_dbg + ((blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 2 + (threadIdx.x != 0)) * 16; \
    for (int i__ = 0; i__ < 8; ++i__) o__[i__] = fd_acc[i__]; o__[8] = (iters); o__[9] = clock64() - fd_t0; \
    long long gt__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__)); o__[10] = gt__; } } while (0)
// backward kernel A: record r of (thread == tid) lands in row 3 * cta + r
#define FDBG_DUMP_ROW(iters, tid, r) do { if (blockIdx.x < 20 && threadIdx.x == (tid)) { \
    long long* o__ = g_flash_dbg + (blockIdx.x * 3 + (r)) * 16; \
    for (int i__ = 0; i__ < 8; ++i__) o__[i__] = fd_acc[i__]; o__[8] = (iters); o__[9] = clock64() - fd_t0; \
    for (int i__ = 8; i__ < 12; ++i__) o__[i__ + 3] = fd_acc[i__]; \
    long long gt__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__)); o__[10] = gt__; } } while (0)
#else
#define FDBG_DECL
#define FDBG(slot)
#define FDBG_DUMP(iters)
#define FDBG_DUMP_ROW(iters, tid, r)
#endif

struct Strides {
    long long bs;
    int rs;
};

// ================================================================================================
// forward
// ================================================================================================
// Division by a launch-time constant as multiply-high + shift (exact for 0 <= n < 2^31): the persistent kernels
// decode an item index in every role warp at every item change, and a hardware-less integer division is ~20 instructions
// on the issue slots the softmax warps need.
struct FastDiv {
    uint32_t mul, shr, d;
    __device__ __forceinline__ int div(int n) const {
        return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> shr);
    }
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = static_cast<uint32_t>(d);
    f.mul = 0;
    f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1ll << lg) < d) ++lg;   // ceil(log2 d)
        const int pw = 31 + lg;
        f.mul = static_cast<uint32_t>(((1ull << pw) + static_cast<uint64_t>(d) - 1) / static_cast<uint64_t>(d));
        f.shr = static_cast<uint32_t>(pw - 32);
    }
    return f;
}
// Items are dealt to the CTAs in boustrophedon order (round r: CTA x takes item r * grid + x, round r + 1: (r + 1) * grid +
// grid - 1 - x): the items are sorted heaviest first, so plain striding gives CTA 0 the heaviest item of EVERY round
// (backward, T = 1024: 51 vs 46.7 128-query blocks on average, 48 with the snake; forward with two CTAs per SM: 27 vs
// 23.4, 24 with the snake).
__device__ __forceinline__ int bwd_first_item() { return static_cast<int>(blockIdx.x); }
__device__ __forceinline__ int bwd_next_item(int item, const FastDiv& div_grid) {
    const int g = static_cast<int>(gridDim.x);
    if (item < 0) return bwd_first_item();   // the "before the first item" state of a cursor
    const int r = div_grid.div(item), x = item - r * g;
    return (r + 1) * g + (g - 1 - x);
}
struct FlashFwdParams {
    bf16* o;
    float* lse;
    Strides os;
    int B, H, Tq, Tk, causal;
    int q_rows;        // rows [0, q_rows) of every (b, h) are produced here (whole 128-row blocks + a partial last one)
    int nqb;           // query blocks per (b, h) = ceil(q_rows / 128)
    int wide_store;    // output rows are 32-byte aligned: 256-bit stores
    float scale, scale_log2e;
    FastDiv div_hb, div_h, div_nqb, div_grid;   // by H * B, by H, by nqb, by the grid size
};

// One step of a CTA's flat sequence of (work item, key block) iterations.  A work item is one 128-row query block of one
// (batch, head); a persistent CTA walks items  blockIdx.x, blockIdx.x + gridDim.x, ...
struct FwdIter {
    int n;       // ordinal of the item inside this CTA (Q slot = n & 1)
    int item;    // global item index; < 0 = past the end
    int j, nkb;  // key block, number of key blocks the item sees
    int q0, h, b;
};
__device__ __forceinline__ void fwd_item_setup(FwdIter& it, const FlashFwdParams& p, int item, int nkb_all) {
    it.item = item;
    int qb, r;
    if (p.causal) {   // heaviest query blocks first: the per-CTA item lists then end with the cheapest items
        const int qi = p.div_hb.div(item);
        qb = p.nqb - 1 - qi;
        r = item - qi * (p.H * p.B);
    } else {          // query blocks of one (b, h) next to each other (their K / V meet in L2)
        r = p.div_nqb.div(item);
        qb = item - r * p.nqb;
    }
    it.b = p.div_h.div(r);
    it.h = r - it.b * p.H;
    it.q0 = qb * BQ;
    it.nkb = p.causal ? min(nkb_all, (it.q0 + BQ - 1 + (p.Tk - p.Tq)) / BK + 1) : nkb_all;
    it.j = 0;
}
__device__ __forceinline__ bool fwd_advance(FwdIter& it, const FlashFwdParams& p, int num_items, int nkb_all) {
    if (it.item < 0) return false;
    if (++it.j < it.nkb) return true;
    const int next = bwd_next_item(it.item, p.div_grid);
    if (next >= num_items) {
        it.item = -1;
        return false;
    }
    ++it.n;
    fwd_item_setup(it, p, next, nkb_all);
    return true;
}

// Streaming attention forward, persistent and warp-specialised.  What an iteration costs is instruction issue and the
// latency chain QK -> softmax -> PV (ncu of the first tcgen05 version: 35 % issue-active, the rest fixed-latency and
// scoreboard waits with two warps per scheduler; phase counters of the bring-up build: profiles/r02/flash_fwd_phases.md),
// so the loop is built to touch every score ONCE, to keep both tensor products, all issue duties and every load latency
// off the softmax warps' critical path, and to have four softmax warps per scheduler:
//   * persistent: 2 CTAs per SM walk the work items (128-row query block of one (b, h)); the flat sequence of
//     (item, key block) iterations is ONE software pipeline — the next item's Q (double-buffered), K and V are in flight
//     and its first QK product is issued while the softmax warps still normalise and store the current item's output
//     (a non-persistent CTA spent ~3,300 clk of prologue per item against ~3,000 per iteration);
//   * warps 0..7 = softmax: a query row is shared by two threads (warp w and w + 4 own the same 32 TMEM lanes), each
//     holding 64 of the row's 128 scores in registers after ONE tcgen05.ld pass (max, then exp2 from the same registers);
//     the two partial row maxima meet in shared memory under a 64-thread named barrier;
//   * warps 8..11 = PV issue | QK issue | K and Q loads | V loads (each converged, one elected lane, its own cursor):
//     QK(g+1) is issued as soon as the eight softmax warps have S(g) in registers (mbarrier s_free) and runs under the
//     exponentials of iteration g; PV(g) is issued when P(g) is complete (mbarrier p_ready) and runs under the loads
//     of iteration g+1;
//   * O accumulates in tensor memory across key blocks (tcgen05.mma accumulate), it is NOT read out every iteration;
//     the running maximum is only raised when it grows by more than 2^8 (a stale maximum is exact arithmetic: P and the
//     row sum are scaled by the same factor, bounded by 256), and only then is O rescaled in tensor memory;
//   * S, P and O have their own columns: S fp32 [0,128) | P bf16 [128,192) | O fp32 [192,256);
//   * K and V have separate double buffers and barriers: K(g+2) is fetched once QK(g) has completed, V(g+1) once PV(g-1) has;
//   * 32-column chunks that no row of the warp can see (causal diagonal) skip their exponentials;
//   * the last key block may be narrower than 128 (a multiple of 16 columns: CLIP's 257th key costs one 16-wide MMA).
// 256 TMEM columns and ~100 KB of shared memory per CTA: two CTAs per SM overlap each other's phases.
// smem: Q0 | Q1 | K0 | K1 | V0 | V1 | row-max exchange [2][2][128] + row-sum exchange [2][128] | barriers
constexpr int kFwdSoftmaxWarps = 8;
// Three warpgroups: two softmax warpgroups and one whose first warp issues TMA / MMA (the other three only give their
// registers away).  Registers are allocated per warpgroup, so 9 warps would cost 12 anyway: 2 CTAs x 384 threads leave
// 80 registers per thread at launch; setmaxnreg moves the issue warpgroup's surplus to the softmax warpgroups
// (2 x 128 x 104 + 128 x 32 = 384 x 80: setmaxnreg only moves registers INSIDE the allocation of the launch — asking for more hangs the CTA in setmaxnreg.inc).
constexpr int kFwdThreads = 384;
constexpr int kFwdSoftmaxRegs = 104, kFwdIssueRegs = 32;
constexpr int kFwdXchBytes = 4096;
__global__ void __launch_bounds__(kFwdThreads, 2)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, FlashFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;               // [2]
    uint8_t* sK = smem + 2 * kTile;   // [2]
    uint8_t* sV = smem + 4 * kTile;   // [2]
    float* sX = reinterpret_cast<float*>(smem + 6 * kTile);   // maxima, then sums: each [2 parities][2 halves][128 rows]
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + 6 * kTile + kFwdXchBytes);   // [2]
    uint64_t* bar_k = bar_q + 2;   // [2]
    uint64_t* bar_v = bar_q + 4;   // [2]
    uint64_t* bar_s = bar_q + 6;        // MMA -> softmax: S(g) complete
    uint64_t* bar_o = bar_q + 7;        // MMA -> softmax: PV(g) complete
    uint64_t* bar_sfree = bar_q + 8;    // softmax (8 warps) -> MMA: S(g) is in registers
    uint64_t* bar_p = bar_q + 9;        // softmax (8 warps) -> MMA: P(g) (and a rescaled O) are in tensor memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 10);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int num_items = p.nqb * p.H * p.B;
    const int nkb_all = (p.Tk + BK - 1) / BK;
    const int n_last = ((p.Tk - (nkb_all - 1) * BK) + 15) & ~15;   // columns of the last key block of the sequence

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        for (int i = 0; i < 8; ++i) ptx::mbar_init(bar_q + i, 1);
        ptx::mbar_init(bar_sfree, kFwdSoftmaxWarps);
        ptx::mbar_init(bar_p, kFwdSoftmaxWarps);
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
    const uint32_t tS = tmem, tP = tmem + 128, tO = tmem + 192;

    // (setmaxnreg sits INSIDE each role's branch: ptxas sizes a region's registers by the setmaxnreg that dominates it —
    // placed before the branch, the softmax code was compiled for the 80 registers of the launch and spilled)
    if (warp >= kFwdSoftmaxWarps) {
        // ===================================== TMA and MMA issue: four warps, one duty each =====
        // PV MMAs | QK MMAs | TMA of K and Q | TMA of V, each a converged warp (one elected lane issues) with its OWN cursor
        // over the CTA's flat sequence of (item, key block) iterations and no ordering between them except the barriers.
        // As one in-order warp with five cursors at 32 registers the issue side spilled, and QK(g+1) — released by
        // "S(g) is in registers" — queued behind the wait for "P(g) is written" of the PV it had to issue first.
        ptx::setmaxnreg_dec<kFwdIssueRegs>();
        const bool issuer = ptx::elect_one();
        const int role = warp - kFwdSoftmaxWarps;
        FwdIter c;
        c.n = 0;
        fwd_item_setup(c, p, blockIdx.x, nkb_all);
        bool valid = true;
        if (role == 0) {
            // ---- O += P(g) V: released by "P(g) (and a rescaled O) are in tensor memory" ----
            const uint64_t dV0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sV), 8192, 1024);   // V [key x 64] read MN-major
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);
#pragma unroll 1
            for (int g = 0; valid; ++g) {
                const int bc = (c.j == nkb_all - 1) ? n_last : BK;
                ptx::mbar_wait(bar_p, g & 1);
                ptx::mbar_wait(&bar_v[g & 1], (g >> 1) & 1);
                ptx::tc_fence_after_sync();
                if (issuer) {
                    const uint64_t dv = dV0 + static_cast<uint32_t>(((g & 1) * kTile) >> 4);
                    const int ksteps = bc >> 4;
                    if (ksteps == 8) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) ptx::umma_bf16_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc, (c.j | k) != 0);
                    } else {
                        for (int k = 0; k < ksteps; ++k)
                            ptx::umma_bf16_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc, (c.j | k) != 0);
                    }
                    ptx::umma_commit(bar_o);
                }
                __syncwarp();
                valid = fwd_advance(c, p, num_items, nkb_all);
            }
        } else if (role == 1) {
            // ---- S(g) = Q K^T (N = the block's columns): released by "S(g-1) is in registers" ----
            const uint64_t dQ0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ), 16, 1024);
            const uint64_t dK0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sK), 16, 1024);
#pragma unroll 1
            for (int g = 0; valid; ++g) {
                const int bc = (c.j == nkb_all - 1) ? n_last : BK;
                if (g >= 1) ptx::mbar_wait(bar_sfree, (g - 1) & 1);
                ptx::mbar_wait(&bar_k[g & 1], (g >> 1) & 1);
                if (c.j == 0) ptx::mbar_wait(&bar_q[c.n & 1], (c.n >> 1) & 1);
                ptx::tc_fence_after_sync();
                if (issuer) {
                    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, bc, 0, 0);
                    const uint64_t dq = dQ0 + static_cast<uint32_t>(((c.n & 1) * kTile) >> 4);
                    const uint64_t dk = dK0 + static_cast<uint32_t>(((g & 1) * kTile) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss(tS, dq + k * 2, dk + k * 2, idesc, k != 0);
                    ptx::umma_commit(bar_s);
                }
                __syncwarp();
                valid = fwd_advance(c, p, num_items, nkb_all);
            }
        } else if (role == 2) {
            // ---- K(g) into buffer g & 1 once S(g-2) was read (so QK(g-2) has completed); an item's Q with its first K:
            //      the Q slot's previous tenant (item n - 2) finished its last QK at iteration g - 2 at the latest ----
#pragma unroll 1
            for (int g = 0; valid; ++g) {
                if (g >= 2) ptx::mbar_wait(bar_sfree, (g - 2) & 1);
                const int c_h = __shfl_sync(0xffffffffu, c.h * 64, 0), c_b = __shfl_sync(0xffffffffu, c.b, 0);
                const int c_k = __shfl_sync(0xffffffffu, c.j * BK, 0), c_q = __shfl_sync(0xffffffffu, c.q0, 0);
                if (issuer) {
                    ptx::mbar_arrive_expect_tx(&bar_k[g & 1], kTile);
                    ptx::tma_load_3d(sK + (g & 1) * kTile, &tmap_k, &bar_k[g & 1], c_h, c_k, c_b);
                    if (c.j == 0) {
                        ptx::mbar_arrive_expect_tx(&bar_q[c.n & 1], kTile);
                        ptx::tma_load_3d(sQ + (c.n & 1) * kTile, &tmap_q, &bar_q[c.n & 1], c_h, c_q, c_b);
                    }
                }
                __syncwarp();
                valid = fwd_advance(c, p, num_items, nkb_all);
            }
        } else {
            // ---- V(g) into buffer g & 1 once PV(g-2) has completed ----
#pragma unroll 1
            for (int g = 0; valid; ++g) {
                if (g >= 2) ptx::mbar_wait(bar_o, (g - 2) & 1);
                const int c_h = __shfl_sync(0xffffffffu, c.h * 64, 0), c_b = __shfl_sync(0xffffffffu, c.b, 0);
                const int c_k = __shfl_sync(0xffffffffu, c.j * BK, 0);
                if (issuer) {
                    ptx::mbar_arrive_expect_tx(&bar_v[g & 1], kTile);
                    ptx::tma_load_3d(sV + (g & 1) * kTile, &tmap_v, &bar_v[g & 1], c_h, c_k, c_b);
                }
                __syncwarp();
                valid = fwd_advance(c, p, num_items, nkb_all);
            }
        }
    } else {
        // ===================================== softmax ==========================================
        ptx::setmaxnreg_inc<kFwdSoftmaxRegs>();
        const int half = warp >> 2;                 // which 64 of the block's 128 score columns (32 of the 64 O columns)
        const int row = ((warp & 3) << 5) + lane;   // query row inside the block = TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t pair_bar = 1 + (warp & 3);   // named barrier of the two warps that share these rows
        const float sl2e = p.scale_log2e;
        const int col_h = half * 64;   // this thread's first score column inside the block
        const int shift = p.Tk - p.Tq;
        FDBG_DECL;
        int g = 0;
        FwdIter it;
        it.n = 0;
        for (int item = blockIdx.x; item < num_items; item = bwd_next_item(item, p.div_grid), ++it.n) {
            fwd_item_setup(it, p, item, nkb_all);
            const int nkb = it.nkb;
            const int qi = it.q0 + row;
            int lim = p.Tk;  // keys [0, lim) visible to this row
            if (p.causal) lim = min(p.Tk, qi + shift + 1);
            if (lim < 1) lim = 1;
            const int lim_min = __reduce_min_sync(0xffffffffu, lim), lim_max = __reduce_max_sync(0xffffffffu, lim);
            float m_ref = -INFINITY, l0 = 0.f, l1 = 0.f;

            for (int j = 0; j < nkb; ++j, ++g) {
                const int bc = (j == nkb_all - 1) ? n_last : BK;
                const int k0 = j * BK;
                FDBG(7);
                ptx::mbar_wait(bar_s, g & 1);
                ptx::tc_fence_after_sync();
                FDBG(0);
                uint32_t s[64];
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (col_h + c * 32 < bc)
                        ptx::tmem_ld_32x32b_x32(tS + lane_base + col_h + c * 32,
                                                *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
                ptx::tmem_ld_wait();
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_sfree);   // this warp's scores are in registers
                FDBG(1);
                // ---- this thread's half of the row maximum (from registers) ----
                // masked: some (row, key) pair of this block is invisible to this warp, or the block is narrower than 128
                const bool masked = (k0 + bc > lim_min) || (bc < BK);
                float mj = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int kc = k0 + col_h + c * 32;
                    if (col_h + c * 32 < bc && kc < lim_max) {   // warp-uniform
                        if (!masked) {
#pragma unroll
                            for (int i = 0; i < 32; i += 2)
                                mj = fmaxf(mj, fmaxf(__uint_as_float(s[c * 32 + i]), __uint_as_float(s[c * 32 + i + 1])));
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (kc + i < lim) mj = fmaxf(mj, __uint_as_float(s[c * 32 + i]));
                        }
                    }
                }
                float* xch = sX + (g & 1) * 256;
                xch[half * 128 + row] = mj;
                ptx::named_bar_sync(pair_bar, 64);
                FDBG(2);
                mj = fmaxf(mj, xch[(half ^ 1) * 128 + row]);
                const float m_new = fmaxf(m_ref, mj);
                // raise the reference maximum only when it would grow by more than 2^8 (exact: P and the sum share the factor)
                const bool raise = (m_new - m_ref) * sl2e > 8.0f;   // -inf -> finite: true; -inf -> -inf: NaN, false
                FDBG(3);
                if (j > 0) {   // PV(g-1) must be complete before P is overwritten or O is rescaled
                    ptx::mbar_wait(bar_o, (g - 1) & 1);
                    ptx::tc_fence_after_sync();
                }
                if (__any_sync(0xffffffffu, raise)) {
                    const float corr = raise ? (m_ref == -INFINITY ? 0.f : ex2f((m_ref - m_new) * sl2e)) : 1.0f;
                    l0 *= corr;
                    l1 *= corr;
                    if (raise) m_ref = m_new;
                    if (j > 0) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32b_x32(tO + lane_base + half * 32, r);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
                        ptx::tmem_st_32x32b_x32(tO + lane_base + half * 32, r);
                    }
                }
                FDBG(4);
                const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * sl2e;
                // ---- probabilities (bf16) -> their own tensor-memory columns ----
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int kc = k0 + col_h + c * 32;
                    if (col_h + c * 32 < bc) {
                        uint32_t pk[16];
                        if (kc < lim_max) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float x0, x1;
                                ptx::ffma2_bcast(x0, x1, __uint_as_float(s[c * 32 + 2 * i]),
                                                 __uint_as_float(s[c * 32 + 2 * i + 1]), sl2e, -mb);
                                float e0 = ex2f(x0), e1 = ex2f(x1);
                                if (masked) {
                                    if (kc + 2 * i >= lim) e0 = 0.f;
                                    if (kc + 2 * i + 1 >= lim) e1 = 0.f;
                                }
                                ptx::fadd2_acc(l0, l1, e0, e1);
                                const bf162 h2 = __floats2bfloat162_rn(e0, e1);
                                pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                            }
                        } else {   // no row of this warp sees these 32 keys (causal diagonal)
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = 0u;
                        }
                        ptx::tmem_st_32x32b_x16(tP + lane_base + (col_h >> 1) + c * 16, pk);
                    }
                }
                FDBG(5);
                ptx::tmem_st_wait();
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar_p);   // this warp's part of P(g) (and of a rescaled O) is in tensor memory
                FDBG(6);
            }
            // ---- the item's output: O / row sum.  The two halves of the row sum meet in shared memory. ----
            float* xs = sX + 512 + (it.n & 1) * 256;
            xs[half * 128 + row] = l0 + l1;
            ptx::mbar_wait(bar_o, (g - 1) & 1);
            ptx::tc_fence_after_sync();
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tO + lane_base + half * 32, r);
            ptx::named_bar_sync(pair_bar, 64);
            const float l = (l0 + l1) + xs[(half ^ 1) * 128 + row];
            ptx::tmem_ld_wait();
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            bf16* orow = p.o + it.b * p.os.bs + static_cast<size_t>(qi) * p.os.rs + it.h * 64 + half * 32;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const bf162 h2 = __floats2bfloat162_rn(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv);
                pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            if (qi < p.q_rows) {
                if (p.wide_store) {
                    ptx::stg_v8(orow, pk);
                    ptx::stg_v8(orow + 16, pk + 8);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        stg16(orow + q * 8, make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]));
                }
                if (half == 0 && p.lse != nullptr)
                    p.lse[(static_cast<size_t>(it.b) * p.H + it.h) * p.Tq + qi] = m_ref * p.scale + __logf(l);
            }
            FDBG(7);
        }
        FDBG_DUMP(g);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, 256);
    }
}

// ================================================================================================
// backward
// ================================================================================================
// delta[b,h,i] = dO_i . O_i
__global__ void __launch_bounds__(128)
flash_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta, int H, int Tq,
                   Strides os, int total_rows) {
    // 8 lanes per (b,h,row): 16-byte loads, 4 rows per warp, 16 rows per block
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const int w = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    float s = 0.f;
    if (w < total_rows) {
        const int qi = w % Tq, h = (w / Tq) % H, b = w / (Tq * H);
        const size_t off = b * os.bs + static_cast<size_t>(qi) * os.rs + h * 64 + sub * 8;
        float a[8], g[8];
        unpack8(ldg16(o + off), a);
        unpack8(ldg16(d_o + off), g);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(a[i], g[i], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0 && w < total_rows) delta[w] = s;
}

struct FlashBwdParams {
    const float* lse;
    const float* delta;
    bf16* out0;  // dK (kernel A) or dQ (kernel B)
    bf16* out1;  // dV (kernel A)
    Strides s0, s1;
    int H, Tq, Tk, causal;
    float scale, scale_log2e;
    FastDiv div_hb, div_h, div_grid;   // by H * B, by H, by the grid size (persistent backward kernels)
};

// ---- kernel A: dK_j, dV_j — persistent, warp-specialised, TWO 64-query sub-blocks in flight (one CTA per SM) ----
// Work item = one 128-key block of one (batch, head); the CTA walks the 64-query sub-blocks that see it.  Everything is
// computed TRANSPOSED so that the TMEM-resident operand is always the A operand (thread = key row = TMEM lane):
//     S^T = K Q_u^T,  dP^T = V dO_u^T   ->  P^T = exp2(S^T * scale*log2e - lse_u*log2e),  dS^T = P^T (dP^T - delta_u)
//     dV += P^T dO_u,  dK += dS^T Q_u     (accumulated IN tensor memory over the sub-blocks, read out once per item)
// TMEM (all 512 columns), slot s = sub-iteration parity:  S^T [128 s, +64) | dP^T [128 s + 64, +64) | P^T bf16
// [256 + 64 s, +32) | dS^T bf16 [288 + 64 s, +32) | dV fp32 [384, 448) | dK fp32 [448, 512).
//   * two softmax groups of eight warps, group g owns slot g: a key row is shared by two threads of the group (32 query
//     columns each), ONE tcgen05.ld pass over S^T and dP^T, probabilities and score gradients go back as bf16.  The two
//     groups are half an iteration out of phase, so one group's tcgen05.ld / barrier / statistics latencies run under the
//     other group's exponentials — a single group of sixteen warps (all waiting on the same barriers) and the two-CTA
//     kernel with per-iteration read-out both measured 138-140 us (profiles/r02/r02_notes.md);
//   * warp 16 (converged, one elected lane issues): when P^T / dS^T of sub-iteration G are complete (p_ready[slot]) it
//     issues their gradient products and, right behind them, the scores of G+2 into the same slot;
//   * Q / dO sub-tiles in a ring of four, K / V in a ring of four items (the load cursor runs three sub-iterations ahead).
constexpr int BQS = 64;                   // queries per sub-block
constexpr int kSubTile = BQS * 128;       // bytes of a 64-row x 64-col bf16 tile
constexpr int kBwdEnd = -0x7fffffff;   // BwdIter::item once the CTA's items are exhausted
struct BwdIter {
    int n, item;      // ordinal of the (non-empty) item inside this CTA (K/V slot = n % 4), global item index (< 0: end)
    int u, u0, nsub;  // 64-query sub-block, first sub-block that sees the key block, sub-blocks of the sequence
    int k0, h, b;
};
__device__ __forceinline__ void bwd_item_setup(BwdIter& it, const FlashBwdParams& p, int item, int B) {
    it.item = item;
    const int hb = p.H * B;
    const int kb = p.div_hb.div(item), r = item - kb * hb;   // key block 0 sees the most query blocks: heaviest items first
    it.b = p.div_h.div(r);
    it.h = r - it.b * p.H;
    it.k0 = kb * BK;
    // whole 128-query blocks, i.e. an EVEN number of sub-blocks per item (a sub-block past the sequence or below the
    // causal diagonal is masked to zero): every item starts in slot 0, both groups do the same number of sub-iterations,
    // and an item never has fewer than two — which the K / V ring of three relies on
    it.nsub = 2 * ((p.Tq + 2 * BQS - 1) / (2 * BQS));
    it.u0 = p.causal ? 2 * (max(0, it.k0 - (p.Tk - p.Tq)) / (2 * BQS)) : 0;
    it.u = it.u0;
}
__device__ __forceinline__ bool bwd_advance(BwdIter& it, const FlashBwdParams& p, int num_items, int B) {
    if (it.item == kBwdEnd) return false;
    if (++it.u < it.nsub) return true;
    for (int next = bwd_next_item(it.item, p.div_grid); next < num_items; next = bwd_next_item(next, p.div_grid)) {
        bwd_item_setup(it, p, next, B);
        if (it.u0 < it.nsub) {   // (items no query sees are handled by the softmax warps alone and are not counted)
            ++it.n;
            return true;
        }
    }
    it.item = kBwdEnd;
    return false;
}

constexpr int kDkvSoftmaxWarps = 16;                       // two groups of eight
constexpr int kDkvThreads = (kDkvSoftmaxWarps + 4) * 32;   // + the issue warpgroup (one active warp)
constexpr int kDkvSoftmaxRegs = 104, kDkvIssueRegs = 64;   // 512 x 104 + 128 x 64 <= 640 x 96 (112 + 64 exceeds the 640 x 96 of the launch: the CTA hangs in setmaxnreg.inc)
// smem: (K, V) x 3 | (Q, dO) sub-tiles x 6 | statistics [2 groups][2][128] | barriers
constexpr int kNKV = 3, kNQ = 6;   // K / V ring (items), Q / dO ring (sub-blocks)
constexpr int kDkvOffQ = kNKV * 2 * kTile, kDkvOffStat = kDkvOffQ + kNQ * 2 * kSubTile, kDkvOffBar = kDkvOffStat + 2 * 2 * 128 * 4;
constexpr int kDkvSmem = kDkvOffBar + 512 + 1024;

__global__ void __launch_bounds__(kDkvThreads, 1)
flash_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     FlashBwdParams p, int B) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem;                    // [kNKV][K | V]
    uint8_t* sQdO = smem + kDkvOffQ;        // [kNQ][Q | dO], 64 rows each
    float* sStat = reinterpret_cast<float*>(smem + kDkvOffStat);
    uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + kDkvOffBar);   // [kNKV] TMA -> MMA
    uint64_t* bar_q = bar_kv + kNKV;           // [kNQ] TMA -> MMA
    uint64_t* bar_s = bar_q + kNQ;             // [2] MMA -> group: S^T, dP^T of the slot complete
    uint64_t* bar_acc = bar_s + 2;             // [2] MMA -> group / issue warp: the gradient products of the slot have retired
    uint64_t* bar_p = bar_acc + 2;             // [2] group (8 warps) -> MMA: P^T, dS^T of the slot are in tensor memory
    uint64_t* bar_sfree = bar_p + 2;           // [2] group (8 warps) -> MMA: S^T, dP^T of the slot are in registers
    uint64_t* bar_done = bar_sfree + 2;        // all 16 warps -> MMA: dV / dK of the finished item have been read out
    uint64_t* bar_item = bar_done + 1;         // MMA -> all 16 warps: every gradient product of the item has retired (one phase
                                               // per item, waited for by both groups in order: a group never polls the OTHER
                                               // slot's bar_acc, whose phase it does not track)
    uint64_t* bar_free = bar_item + 1;         // [kNQ] MMA -> TMA: the products that read the Q / dO ring slot have retired
    uint64_t* bar_kvfree = bar_free + kNQ;     // [kNKV] MMA -> TMA: every product of the item in the K / V slot has retired
    uint64_t* bar_stat = bar_kvfree + kNKV;    // [2 groups][2 buffers] 4 writer warps -> the group: statistics are staged
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_stat + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int nkb = (p.Tk + BK - 1) / BK;
    const int num_items = nkb * p.H * B;
    const int shift = p.Tk - p.Tq;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        ptx::prefetch_tensormap(&tmap_do);
        for (int i = 0; i < kNKV + kNQ + 4; ++i) ptx::mbar_init(bar_kv + i, 1);
        ptx::mbar_init(bar_item, 1);
        for (int i = 0; i < kNQ + kNKV; ++i) ptx::mbar_init(bar_free + i, 1);
        for (int i = 0; i < 4; ++i) ptx::mbar_init(bar_stat + i, 4);
        for (int i = 0; i < 4; ++i) ptx::mbar_init(bar_p + i, kDkvSoftmaxWarps / 2);   // bar_p[2], bar_sfree[2]
        ptx::mbar_init(bar_done, kDkvSoftmaxWarps);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tDV = tmem + 384, tDK = tmem + 448;

    // ===================================== TMA, score MMAs, product MMAs: one warp each ========================
    // Three converged warps on three different schedulers (one elected lane issues), each with its own cursor over the
    // CTA's sub-iterations and NO ordering between them except the barriers: as one in-order warp the issue side was busy
    // ~1,450 clk per sub-iteration (it shares its scheduler with four softmax warps) and bounded the whole kernel.
    // (setmaxnreg sits INSIDE each role's branch: ptxas sizes a region's registers by the setmaxnreg that dominates it)
    if (warp >= kDkvSoftmaxWarps) {
        ptx::setmaxnreg_dec<kDkvIssueRegs>();
        const bool issuer = ptx::elect_one();
        BwdIter c;   // "before the first item": the first advance finds the first item some query sees
        c.n = -1;
        c.item = -1;
        c.u = c.u0 = c.nsub = 0;
        c.k0 = c.h = c.b = 0;
        bool valid = bwd_advance(c, p, num_items, B);
        if (warp == kDkvSoftmaxWarps + 3) {
            valid = false;   // the warpgroup's fourth warp has no role
        }
        if (warp == kDkvSoftmaxWarps + 2) {
            // ---- TMA: Q / dO of every sub-iteration (ring of six), K / V of every item (ring of three) ----
#pragma unroll 1
            for (int G = 0; valid; ++G) {
                const int r = G % kNQ, kvs = c.n % kNKV;
                const bool first = c.u == c.u0;
                if (G >= kNQ) ptx::mbar_wait(&bar_free[r], (G / kNQ - 1) & 1);
                if (first && c.n >= kNKV) ptx::mbar_wait(&bar_kvfree[kvs], (c.n / kNKV - 1) & 1);
                const int c_h = __shfl_sync(0xffffffffu, c.h * 64, 0), c_b = __shfl_sync(0xffffffffu, c.b, 0);
                const int c_q = __shfl_sync(0xffffffffu, c.u * BQS, 0), c_k = __shfl_sync(0xffffffffu, c.k0, 0);
                if (issuer) {
                    if (first) {
                        uint8_t* dst = sKV + kvs * 2 * kTile;
                        ptx::mbar_arrive_expect_tx(&bar_kv[kvs], 2 * kTile);
                        ptx::tma_load_3d(dst, &tmap_k, &bar_kv[kvs], c_h, c_k, c_b);
                        ptx::tma_load_3d(dst + kTile, &tmap_v, &bar_kv[kvs], c_h, c_k, c_b);
                    }
                    uint8_t* dst = sQdO + r * 2 * kSubTile;
                    ptx::mbar_arrive_expect_tx(&bar_q[r], 2 * kSubTile);
                    ptx::tma_load_3d(dst, &tmap_q, &bar_q[r], c_h, c_q, c_b);
                    ptx::tma_load_3d(dst + kSubTile, &tmap_do, &bar_q[r], c_h, c_q, c_b);
                }
                __syncwarp();
                valid = bwd_advance(c, p, num_items, B);
            }
        } else if (warp == kDkvSoftmaxWarps + 1) {
            // ---- scores of G: S^T = K Q^T, dP^T = V dO^T, as soon as the group has S^T / dP^T of G - 2 in registers ----
            const uint64_t dKV0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sKV), 16, 1024);    // K-major A operand
            const uint64_t dQ0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sQdO), 16, 1024);    // K-major B operand
#pragma unroll 1
            for (int G = 0; valid; ++G) {
                const int kvs = c.n % kNKV;
                if (G >= 2) ptx::mbar_wait(&bar_sfree[G & 1], ((G - 2) >> 1) & 1);
                ptx::mbar_wait(&bar_q[G % kNQ], (G / kNQ) & 1);
                ptx::mbar_wait(&bar_kv[kvs], (c.n / kNKV) & 1);
                ptx::tc_fence_after_sync();
                if (issuer) {
                    const uint64_t ak = dKV0 + static_cast<uint32_t>((kvs * 2 * kTile) >> 4), av = ak + (kTile >> 4);
                    const uint64_t bq = dQ0 + static_cast<uint32_t>(((G % kNQ) * 2 * kSubTile) >> 4), bdo = bq + (kSubTile >> 4);
                    const uint32_t tS = tmem + (G & 1) * 128, tDP = tS + 64;
                    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, BQS, 0, 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ptx::umma_bf16_ss(tS, ak + k * 2, bq + k * 2, idesc, k != 0);
                        ptx::umma_bf16_ss(tDP, av + k * 2, bdo + k * 2, idesc, k != 0);
                    }
                    ptx::umma_commit(&bar_s[G & 1]);
                }
                __syncwarp();
                valid = bwd_advance(c, p, num_items, B);
            }
        } else {
            // ---- gradient products of G: dV += P^T dO, dK += dS^T Q, once the group has written P^T / dS^T ----
            const uint64_t dQmn0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sQdO), 8192, 1024);   // the tiles read MN-major
#pragma unroll 1
            for (int G = 0; valid; ++G) {
                const int slot = G & 1;
                const bool first = c.u == c.u0, last = c.u + 1 == c.nsub;
                ptx::mbar_wait(&bar_p[slot], (G >> 1) & 1);
                if (first && c.n > 0) ptx::mbar_wait(bar_done, (c.n - 1) & 1);   // dV / dK of the previous item were read out
                ptx::tc_fence_after_sync();
                if (issuer) {
                    const uint64_t bdq = dQmn0 + static_cast<uint32_t>(((G % kNQ) * 2 * kSubTile) >> 4);
                    const uint64_t bdo = bdq + (kSubTile >> 4);
                    const uint32_t tP = tmem + 256 + slot * 64, tDS = tP + 32;
                    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);   // B = [query x 64] read MN-major
                    const uint32_t acc = first ? 0u : 1u;
#pragma unroll
                    for (int kk = 0; kk < BQS / 16; ++kk) {
                        ptx::umma_bf16_ts(tDV, tP + kk * 8, bdo + kk * (2048 >> 4), idesc, (acc | kk) != 0);
                        ptx::umma_bf16_ts(tDK, tDS + kk * 8, bdq + kk * (2048 >> 4), idesc, (acc | kk) != 0);
                    }
                    ptx::umma_commit(&bar_acc[slot]);
                    ptx::umma_commit(&bar_free[G % kNQ]);
                    if (last) {
                        ptx::umma_commit(bar_item);
                        ptx::umma_commit(&bar_kvfree[c.n % kNKV]);
                    }
                }
                __syncwarp();
                valid = bwd_advance(c, p, num_items, B);
            }
        }
    } else {
        // ===================================== softmax / gradient of the scores =================
        ptx::setmaxnreg_inc<kDkvSoftmaxRegs>();
        const int grp = warp >> 3;                           // the slot this group owns
        const int w8 = warp & 7;
        const int half = w8 >> 2;                            // which 32 of the sub-block's 64 query columns
        const int row = ((w8 & 3) << 5) + lane;              // key row inside the block = TMEM lane
        const int gt = threadIdx.x & 255;                    // thread index inside the group
        const uint32_t lane_base = static_cast<uint32_t>((w8 & 3) * 32) << 16;
        const int col_h = half * 32;
        const float sl2e = p.scale_log2e;
        const uint32_t tS = tmem + grp * 128 + lane_base + col_h, tDP = tS + 64;
        const uint32_t tP = tmem + 256 + grp * 64 + lane_base + (col_h >> 1), tDS = tP + 32;
        float* stat = sStat + grp * 256;                     // [2 buffers][64 queries x (lse2, lse2', delta, delta') pairs]
        int kown = 0;                                        // this group's sub-iterations so far (phase of its slot's barriers)
        int n_done = 0;                                      // non-empty items finished (phase of bar_item)
        float cur_stat = 0.f;                                // this thread's staged statistic (prefetched one sub-iteration ahead)
        bool have_stat = false;
        BwdIter it;
        FDBG_DECL;
        int my_iters = 0;
        bool pending = false;        // an item whose dV / dK have not been read out yet
        bf16* rout_prev = nullptr;
        int kj_prev = 0;
        auto read_out = [&](bf16* rout_, int kj_) {
            // ---- the item's gradients, once every product of the item has retired ----
            FDBG(7);
            ptx::mbar_wait(bar_item, n_done & 1);
            FDBG(6);
            ++n_done;
            ptx::tc_fence_after_sync();
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32((grp == 0 ? tDV : tDK) + lane_base + half * 32, r);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_done);   // the next item's first products may overwrite dV / dK
            FDBG(5);
            // Four lanes transpose their 4 x 16-byte chunks so that one store instruction writes 64 contiguous bytes per
            // row for eight rows (lane = row would touch 32 different 128-byte lines per instruction: measured 3,100 clk
            // per item in the store queue).
            {
                const float sc = grp == 0 ? 1.0f : p.scale;
                uint32_t w[16];
#pragma unroll
                for (int x = 0; x < 16; ++x) {
                    const bf162 hv = __floats2bfloat162_rn(__uint_as_float(r[2 * x]) * sc, __uint_as_float(r[2 * x + 1]) * sc);
                    w[x] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                const int j = lane & 3;
#pragma unroll
                for (int step = 1; step <= 2; ++step) {   // butterfly: swap chunk c of lane i with chunk c^step of lane i^step
                    const bool hi = (j & step) != 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c & step) continue;            // pair (c, c + step)
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const uint32_t send = hi ? w[4 * c + x] : w[4 * (c + step) + x];
                            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, step);
                            if (hi) w[4 * c + x] = recv;
                            else w[4 * (c + step) + x] = recv;
                        }
                    }
                }
                // now chunk position c holds columns [8 j, 8 j + 8) of row (row - j + c)
                const size_t rs = grp == 0 ? p.s1.rs : p.s0.rs;
                bf16* rq = rout_ + 8 * j - static_cast<size_t>(j) * rs;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (kj_ - j + c < p.Tk) stg16(rq + c * rs, make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
            }
            FDBG(8);
        };
        for (int item = bwd_first_item();; item = bwd_next_item(item, p.div_grid)) {
            // one pass per item plus a final flush pass (no sub-iteration) that reads out the last item: ONE copy of the
            // read-out code
            const bool live = item < num_items;
            if (!live && !pending) break;
            int kj = 0, u_beg = 0, u_end = 0;
            bf16* rout = nullptr;
            size_t stat_base = 0;
            bool tail_keys = false;
            if (live) {
                bwd_item_setup(it, p, item, B);
                kj = it.k0 + row;
                // read-out assignment: group 0 stores dV, group 1 stores dK; a row is shared by the group's two threads
                rout = (grp == 0 ? p.out1 + it.b * p.s1.bs + static_cast<size_t>(kj) * p.s1.rs
                                 : p.out0 + it.b * p.s0.bs + static_cast<size_t>(kj) * p.s0.rs) + it.h * 64 + half * 32;
                if (it.u0 >= it.nsub) {   // no query sees this key block: zero gradients
                    if (kj < p.Tk) {
                        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int c = 0; c < 32; c += 8) stg16(rout + c, z);
                    }
                    continue;
                }
                stat_base = (static_cast<size_t>(it.b) * p.H + it.h) * p.Tq;
                tail_keys = it.k0 + BK > p.Tk;   // some key rows of this block are past the sequence
                FDBG(9);
                // (an item has an even number of sub-iterations and starts in slot 0: this group's are u0 + grp, + 2, ...)
                u_beg = it.u0 + grp;
                u_end = it.nsub;
            }
            int u = u_beg;
            do {
                if (u < u_end) {
                const int slot = grp;
                const int k = kown;               // the k-th sub-iteration of this slot (barrier phase)
                const int qs = u * BQS;
                ++my_iters;
                FDBG(7);
                // statistics of the sub-block: float4 (-lse * log2e, -lse' * log2e, -delta, -delta') per PAIR of query columns,
                // staged in shared memory by the group's first four warps from values prefetched under the previous
                // sub-iteration's arithmetic; the buffer's mbarrier (four arrivals) replaces a 256-thread named barrier
                // (10 % of the kernel's stall samples).  Staging at the END of the previous sub-iteration instead was
                // measured: 175.6 vs 164.1 us (72 bytes of spills in the arithmetic).  The other buffer was last read in the
                // arithmetic of sub-iteration k - 2, which every warp finished before the products of k - 2 were issued —
                // and this warp waited for those before its tcgen05.st of k - 1.
                float* st = stat + (k & 1) * 128;
                auto load_stat = [&](size_t base_, int q0_) -> float {   // thread gt < 64: lse of query gt; 64 <= gt < 128: delta
                    const int qi = q0_ + (gt & 63);        // (raw values: nothing here may depend on the load, it is in flight)
                    float v = gt < 64 ? INFINITY : 0.f;    // a query past the sequence: p -> 0, dS -> 0
                    if (gt < 128 && qi < p.Tq) v = (gt < 64 ? p.lse : p.delta)[base_ + qi];
                    return v;
                };
                auto stage = [&](float* dst, int buf) {
                    if (gt < 128) {   // (warp-uniform)
                        dst[((gt & 63) >> 1) * 4 + (gt >> 6) * 2 + (gt & 1)] = -cur_stat * (gt < 64 ? 1.4426950408889634f : 1.f);
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&bar_stat[grp * 2 + buf]);
                    }
                };
                if (!have_stat) cur_stat = load_stat(stat_base, qs);
                stage(st, k & 1);
                // this group's next sub-block — of this item or, at its last one, of the CTA's next item — is in flight under
                // the arithmetic
                have_stat = u + 2 < it.nsub;
                if (have_stat) {
                    cur_stat = load_stat(stat_base, qs + 2 * BQS);
                } else {
                    const int nx = bwd_next_item(item, p.div_grid);
                    if (nx < num_items) {
                        BwdIter nt;
                        bwd_item_setup(nt, p, nx, B);
                        if (nt.u0 < nt.nsub) {   // (the next item this CTA processes: empty items have no sub-iteration)
                            cur_stat = load_stat((static_cast<size_t>(nt.b) * p.H + nt.h) * p.Tq, (nt.u0 + grp) * BQS);
                            have_stat = true;
                        }
                    }
                }
                ptx::mbar_wait(&bar_stat[grp * 2 + (k & 1)], (k >> 1) & 1);
                FDBG(0);
                const float4* st4 = reinterpret_cast<const float4*>(st) + (col_h >> 1);
                ptx::mbar_wait(&bar_s[slot], k & 1);
                FDBG(1);
                ptx::tc_fence_after_sync();
                uint32_t s[32], dp[32];
                ptx::tmem_ld_32x32b_x32(tS, s);
                ptx::tmem_ld_32x32b_x32(tDP, dp);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bar_sfree[slot]);   // the scores of G + 2 may overwrite the slot
                FDBG(2);
                const bool slow = tail_keys || (p.causal && (qs + shift < it.k0 + BK - 1));   // masked pairs in this sub-block
                uint32_t pk[16], dk[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float4 sv = st4[t];
                    const float p0 = ex2f(fmaf(__uint_as_float(s[2 * t]), sl2e, sv.x));
                    const float p1 = ex2f(fmaf(__uint_as_float(s[2 * t + 1]), sl2e, sv.y));
                    const float d0 = p0 * (__uint_as_float(dp[2 * t]) + sv.z);
                    const float d1 = p1 * (__uint_as_float(dp[2 * t + 1]) + sv.w);
                    const bf162 hp = __floats2bfloat162_rn(p0, p1);
                    const bf162 hd = __floats2bfloat162_rn(d0, d1);
                    pk[t] = *reinterpret_cast<const uint32_t*>(&hp);
                    dk[t] = *reinterpret_cast<const uint32_t*>(&hd);
                }
                if (slow) {
                    // Masked pairs are cleared in the packed results (bit operations: an overflowed exponential of a masked
                    // score never reaches the products) — one copy of the arithmetic keeps the kernel inside the 32 KB
                    // instruction-cache level.  A query column c of this thread is visible iff c >= cmin: causal = key <=
                    // query + shift; a key past the sequence sees nothing.
                    const int cmin = kj >= p.Tk ? (1 << 30) : (p.causal ? kj - qs - shift - col_h : -(1 << 30));
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        if (2 * t + 1 < cmin) {
                            pk[t] = 0u;
                            dk[t] = 0u;
                        } else if (2 * t < cmin) {
                            pk[t] &= 0xffff0000u;
                            dk[t] &= 0xffff0000u;
                        }
                    }
                }
                FDBG(3);
                if (k > 0) {   // this slot's previous gradient products read P^T / dS^T: retired before the stores
                    ptx::mbar_wait(&bar_acc[slot], (k - 1) & 1);
                    ptx::tc_fence_after_sync();
                }
                ptx::tmem_st_32x32b_x16(tP, pk);
                ptx::tmem_st_32x32b_x16(tDS, dk);
                ptx::tmem_st_wait();
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bar_p[slot]);
                FDBG(4);
                ++kown;
                }
                if (pending) {   // (also the flush pass after the last item, which has no sub-iteration)
                    read_out(rout_prev, kj_prev);
                    pending = false;
                }
                u += 2;
            } while (u < u_end);
            if (!live) break;
            // the read-out of this item is deferred until this group has written the first P^T / dS^T of the NEXT item:
            // the wait for the item's last products (issued for the other group half a period later) runs under useful work
            pending = true;
            rout_prev = rout;
            kj_prev = kj;
        }
        FDBG_DUMP_ROW(my_iters, 0, 0);
        FDBG_DUMP_ROW(my_iters, 256, 1);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, 512);
    }
}

// ---- kernel B: dQ_i — persistent, warp-specialised, TWO 64-key sub-blocks in flight (one CTA per SM) ----
// The same machine as kernel A with the roles of queries and keys exchanged.  Work item = one 128-query block of one
// (batch, head); the CTA walks the 64-key sub-blocks it sees (thread = query row = TMEM lane):
//     S = Q K_v^T,  dP = dO V_v^T   ->   dS = exp2(S * scale*log2e - lse*log2e) (dP - delta)   ->   dQ += dS K_v
// TMEM: slot s = sub-iteration parity:  S [128 s, +64) | dP [128 s + 64, +64) | dS bf16 [256 + 32 s, +32) | dQ fp32
// [320, 384).  Q / dO in a ring of two items, K / V sub-tiles in a ring of six; the row statistics are per thread (two
// loads per item).  Item n is read out by softmax group n & 1 (so the two groups alternate), after that group has written
// the first dS of the next item.
struct DqIter {
    int n, item;      // ordinal of the (non-empty) item inside this CTA, global item index (kBwdEnd: exhausted)
    int v, nsub;      // 64-key sub-block, sub-blocks this query block sees (even; 0 = none)
    int q0, h, b;
};
__device__ __forceinline__ void dq_item_setup(DqIter& it, const FlashBwdParams& p, int item, int B, int nqb) {
    it.item = item;
    const int hb = p.H * B;
    const int qi = p.div_hb.div(item), r = item - qi * hb;
    const int qb = nqb - 1 - qi;   // with a causal mask the last query block sees the most keys: first
    it.b = p.div_h.div(r);
    it.h = r - it.b * p.H;
    it.q0 = qb * BQ;
    const int nkeys = p.causal ? min(p.Tk, it.q0 + BQ + (p.Tk - p.Tq)) : p.Tk;   // keys some row of the block sees
    it.nsub = nkeys > 0 ? 2 * ((nkeys + BK - 1) / BK) : 0;
    it.v = 0;
}
__device__ __forceinline__ bool dq_advance(DqIter& it, const FlashBwdParams& p, int num_items, int B, int nqb) {
    if (it.item == kBwdEnd) return false;
    if (++it.v < it.nsub) return true;
    for (int next = bwd_next_item(it.item, p.div_grid); next < num_items; next = bwd_next_item(next, p.div_grid)) {
        dq_item_setup(it, p, next, B, nqb);
        if (it.nsub > 0) {
            ++it.n;
            return true;
        }
    }
    it.item = kBwdEnd;
    return false;
}
constexpr int kDqNKV = 6;   // K / V sub-tile ring
// smem: (Q, dO) x 2 | (K, V) sub-tiles x 6 | barriers
constexpr int kDqOffKV = 2 * 2 * kTile, kDqOffBar = kDqOffKV + kDqNKV * 2 * kSubTile;
constexpr int kDqSmem = kDqOffBar + 256 + 1024;

__global__ void __launch_bounds__(kDkvThreads, 1)
flash_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                    FlashBwdParams p, int B) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQdO = smem;                  // [2][Q | dO], 128 rows each
    uint8_t* sKV = smem + kDqOffKV;        // [kDqNKV][K | V], 64 rows each
    uint64_t* bar_qdo = reinterpret_cast<uint64_t*>(smem + kDqOffBar);   // [2] TMA -> MMA
    uint64_t* bar_kv = bar_qdo + 2;            // [kDqNKV] TMA -> MMA
    uint64_t* bar_s = bar_kv + kDqNKV;         // [2] MMA -> group: S, dP of the slot complete
    uint64_t* bar_acc = bar_s + 2;             // [2] MMA -> group: the product that read the slot's dS has retired
    uint64_t* bar_p = bar_acc + 2;             // [2] group (8 warps) -> MMA: dS of the slot is in tensor memory
    uint64_t* bar_sfree = bar_p + 2;           // [2] group (8 warps) -> MMA: S, dP of the slot are in registers
    uint64_t* bar_done = bar_sfree + 2;        // [2] group n & 1 (8 warps) -> MMA: dQ of item n has been read out
    uint64_t* bar_item = bar_done + 2;         // [2] MMA -> group n & 1: every product of item n has retired
    uint64_t* bar_free = bar_item + 2;         // [kDqNKV] MMA -> TMA: the K / V ring slot has been consumed
    uint64_t* bar_qfree = bar_free + kDqNKV;   // [2] MMA -> TMA: the item's scores have consumed Q / dO
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qfree + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int nqb = (p.Tq + BQ - 1) / BQ;
    const int num_items = nqb * p.H * B;
    const int shift = p.Tk - p.Tq;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_k);
        ptx::prefetch_tensormap(&tmap_v);
        ptx::prefetch_tensormap(&tmap_do);
        for (int i = 0; i < 2 + kDqNKV + 4; ++i) ptx::mbar_init(bar_qdo + i, 1);          // bar_qdo, bar_kv, bar_s, bar_acc
        for (int i = 0; i < 6; ++i) ptx::mbar_init(bar_p + i, kDkvSoftmaxWarps / 2);      // bar_p, bar_sfree, bar_done
        for (int i = 0; i < 2 + kDqNKV + 2; ++i) ptx::mbar_init(bar_item + i, 1);         // bar_item, bar_free, bar_qfree
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tDQ = tmem + 320;

    if (warp >= kDkvSoftmaxWarps) {
        ptx::setmaxnreg_dec<kDkvIssueRegs>();
        const bool issuer = ptx::elect_one();
        DqIter c;   // "before the first item"
        c.n = -1;
        c.item = -1;
        c.v = c.nsub = 0;
        c.q0 = c.h = c.b = 0;
        bool valid = dq_advance(c, p, num_items, B, nqb);
        if (warp == kDkvSoftmaxWarps + 3) valid = false;   // the warpgroup's fourth warp has no role
        if (warp == kDkvSoftmaxWarps + 2) {
            // ---- TMA: Q / dO of every item (ring of two), K / V of every sub-iteration (ring of six) ----
#pragma unroll 1
            for (int G = 0; valid; ++G) {
                const int r = G % kDqNKV, qs = c.n & 1;
                const bool first = c.v == 0;
                if (G >= kDqNKV) ptx::mbar_wait(&bar_free[r], (G / kDqNKV - 1) & 1);
                if (first && c.n >= 2) ptx::mbar_wait(&bar_qfree[qs], (c.n / 2 - 1) & 1);
                const int c_h = __shfl_sync(0xffffffffu, c.h * 64, 0), c_b = __shfl_sync(0xffffffffu, c.b, 0);
                const int c_q = __shfl_sync(0xffffffffu, c.q0, 0), c_k = __shfl_sync(0xffffffffu, c.v * BQS, 0);
                if (issuer) {
                    if (first) {
                        uint8_t* dst = sQdO + qs * 2 * kTile;
                        ptx::mbar_arrive_expect_tx(&bar_qdo[qs], 2 * kTile);
                        ptx::tma_load_3d(dst, &tmap_q, &bar_qdo[qs], c_h, c_q, c_b);
                        ptx::tma_load_3d(dst + kTile, &tmap_do, &bar_qdo[qs], c_h, c_q, c_b);
                    }
                    uint8_t* dst = sKV + r * 2 * kSubTile;
                    ptx::mbar_arrive_expect_tx(&bar_kv[r], 2 * kSubTile);
                    ptx::tma_load_3d(dst, &tmap_k, &bar_kv[r], c_h, c_k, c_b);
                    ptx::tma_load_3d(dst + kSubTile, &tmap_v, &bar_kv[r], c_h, c_k, c_b);
                }
                __syncwarp();
                valid = dq_advance(c, p, num_items, B, nqb);
            }
        } else if (warp == kDkvSoftmaxWarps + 1) {
            // ---- scores of G: S = Q K^T, dP = dO V^T, as soon as the group has S / dP of G - 2 in registers ----
            const uint64_t dQ0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sQdO), 16, 1024);   // K-major A operand (128 rows)
            const uint64_t dK0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sKV), 16, 1024);    // K-major B operand (64 rows)
#pragma unroll 1
            for (int G = 0; valid; ++G) {
                const int qs = c.n & 1;
                const bool last = c.v + 1 == c.nsub;
                if (G >= 2) ptx::mbar_wait(&bar_sfree[G & 1], ((G - 2) >> 1) & 1);
                ptx::mbar_wait(&bar_kv[G % kDqNKV], (G / kDqNKV) & 1);
                ptx::mbar_wait(&bar_qdo[qs], (c.n >> 1) & 1);
                ptx::tc_fence_after_sync();
                if (issuer) {
                    const uint64_t aq = dQ0 + static_cast<uint32_t>((qs * 2 * kTile) >> 4), ado = aq + (kTile >> 4);
                    const uint64_t bk = dK0 + static_cast<uint32_t>(((G % kDqNKV) * 2 * kSubTile) >> 4), bv = bk + (kSubTile >> 4);
                    const uint32_t tS = tmem + (G & 1) * 128, tDP = tS + 64;
                    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, BQS, 0, 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        ptx::umma_bf16_ss(tS, aq + k * 2, bk + k * 2, idesc, k != 0);
                        ptx::umma_bf16_ss(tDP, ado + k * 2, bv + k * 2, idesc, k != 0);
                    }
                    ptx::umma_commit(&bar_s[G & 1]);
                    if (last) ptx::umma_commit(&bar_qfree[qs]);
                }
                __syncwarp();
                valid = dq_advance(c, p, num_items, B, nqb);
            }
        } else {
            // ---- product of G: dQ += dS K, once the group has written dS ----
            const uint64_t dKmn0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sKV), 8192, 1024);   // K sub-tile read MN-major
#pragma unroll 1
            for (int G = 0; valid; ++G) {
                const int slot = G & 1;
                const bool first = c.v == 0, last = c.v + 1 == c.nsub;
                ptx::mbar_wait(&bar_p[slot], (G >> 1) & 1);
                if (first && c.n > 0) ptx::mbar_wait(&bar_done[(c.n - 1) & 1], ((c.n - 1) >> 1) & 1);   // dQ of n - 1 was read out
                ptx::tc_fence_after_sync();
                if (issuer) {
                    const uint64_t bk = dKmn0 + static_cast<uint32_t>(((G % kDqNKV) * 2 * kSubTile) >> 4);
                    const uint32_t tDS = tmem + 256 + slot * 32;
                    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 1);   // B = [key x 64] read MN-major
                    const uint32_t acc = first ? 0u : 1u;
#pragma unroll
                    for (int kk = 0; kk < BQS / 16; ++kk)
                        ptx::umma_bf16_ts(tDQ, tDS + kk * 8, bk + kk * (2048 >> 4), idesc, (acc | kk) != 0);
                    ptx::umma_commit(&bar_acc[slot]);
                    ptx::umma_commit(&bar_free[G % kDqNKV]);
                    if (last) ptx::umma_commit(&bar_item[c.n & 1]);
                }
                __syncwarp();
                valid = dq_advance(c, p, num_items, B, nqb);
            }
        }
    } else {
        // ===================================== gradient of the scores ===========================
        ptx::setmaxnreg_inc<kDkvSoftmaxRegs>();
        const int grp = warp >> 3;                           // the slot this group owns
        const int w8 = warp & 7;
        const int half = w8 >> 2;                            // which 32 of the sub-block's 64 key columns
        const int row = ((w8 & 3) << 5) + lane;              // query row inside the block = TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>((w8 & 3) * 32) << 16;
        const int col_h = half * 32;
        const float sl2e = p.scale_log2e;
        const uint32_t tS = tmem + grp * 128 + lane_base + col_h, tDP = tS + 64;
        const uint32_t tDS = tmem + 256 + grp * 32 + lane_base + (col_h >> 1);
        int kown = 0;        // this group's sub-iterations so far (phase of its slot's barriers)
        int n = 0;           // ordinal of the current non-empty item
        bool pending = false;        // an item of this group whose dQ has not been read out yet
        bf16* rout_prev = nullptr;
        int qi_prev = 0, n_prev = 0;
        DqIter it;
        FDBG_DECL;
        int my_iters = 0;
        auto read_out = [&](bf16* rout_, int qi_, int n_) {
            ptx::mbar_wait(&bar_item[grp], (n_ >> 1) & 1);
            ptx::tc_fence_after_sync();
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tDQ + lane_base + half * 32, r);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bar_done[grp]);   // the next item's first product may overwrite dQ
            uint32_t w[16];
#pragma unroll
            for (int x = 0; x < 16; ++x) {
                const bf162 hv = __floats2bfloat162_rn(__uint_as_float(r[2 * x]) * p.scale, __uint_as_float(r[2 * x + 1]) * p.scale);
                w[x] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            const int j = lane & 3;   // four lanes transpose their 4 x 16-byte chunks: 64 contiguous bytes per row and store
#pragma unroll
            for (int step = 1; step <= 2; ++step) {
                const bool hi = (j & step) != 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c & step) continue;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const uint32_t send = hi ? w[4 * c + x] : w[4 * (c + step) + x];
                        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, step);
                        if (hi) w[4 * c + x] = recv;
                        else w[4 * (c + step) + x] = recv;
                    }
                }
            }
            const size_t rs = p.s0.rs;
            bf16* rq = rout_ + 8 * j - static_cast<size_t>(j) * rs;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (qi_ - j + c < p.Tq) stg16(rq + c * rs, make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
        };
        for (int item = bwd_first_item();; item = bwd_next_item(item, p.div_grid)) {
            const bool live = item < num_items;
            if (!live && !pending) break;
            int qi = 0, v_beg = 0, v_end = 0;
            bf16* rout = nullptr;
            float lse_raw = INFINITY, delta_raw = 0.f;   // a row past the sequence keeps +inf: its p and dS are zero
            if (live) {
                dq_item_setup(it, p, item, B, nqb);
                qi = it.q0 + row;
                rout = p.out0 + it.b * p.s0.bs + static_cast<size_t>(qi) * p.s0.rs + it.h * 64 + half * 32;
                if (it.nsub == 0) {   // this query block sees no key: zero gradient (group 0 writes it)
                    if (grp == 0 && qi < p.Tq) {
                        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int c = 0; c < 32; c += 8) stg16(rout + c, z);
                    }
                    continue;
                }
                if (qi < p.Tq) {   // nothing here depends on the loaded values: they are first used after the wait for S
                    const size_t si = (static_cast<size_t>(it.b) * p.H + it.h) * p.Tq + qi;
                    lse_raw = p.lse[si];
                    delta_raw = p.delta[si];
                }
                v_beg = grp;   // an item has an even number of sub-iterations and starts in slot 0
                v_end = it.nsub;
            }
            int v = v_beg;
            do {
                if (v < v_end) {
                    const int slot = grp, k = kown;
                    const int k0 = v * BQS;   // first key of the sub-block
                    ++my_iters;
                    FDBG(5);
                    ptx::mbar_wait(&bar_s[slot], k & 1);
                    FDBG(0);
                    ptx::tc_fence_after_sync();
                    uint32_t s[32], dp[32];
                    ptx::tmem_ld_32x32b_x32(tS, s);
                    ptx::tmem_ld_32x32b_x32(tDP, dp);
                    ptx::tmem_ld_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bar_sfree[slot]);   // the scores of G + 2 may overwrite the slot
                    FDBG(1);
                    asm volatile("" : "+f"(lse_raw), "+f"(delta_raw));   // (keeps the scaling below out of the item setup)
                    const float lse2 = -lse_raw * 1.4426950408889634f, ndelta = -delta_raw;
                    uint32_t dk[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const float p0 = ex2f(fmaf(__uint_as_float(s[2 * t]), sl2e, lse2));
                        const float p1 = ex2f(fmaf(__uint_as_float(s[2 * t + 1]), sl2e, lse2));
                        const float d0 = p0 * (__uint_as_float(dp[2 * t]) + ndelta);
                        const float d1 = p1 * (__uint_as_float(dp[2 * t + 1]) + ndelta);
                        const bf162 hd = __floats2bfloat162_rn(d0, d1);
                        dk[t] = *reinterpret_cast<const uint32_t*>(&hd);
                    }
                    // masked pairs (warp-uniform test): a key column c of this thread is visible iff c <= cmax — causal: key
                    // <= query + shift; keys past the sequence (zero-filled rows: their p is NOT zero) see nothing
                    if (k0 + BQS > p.Tk || (p.causal && k0 + BQS - 1 > it.q0 + shift)) {
                        int cmax = p.Tk - 1 - k0;
                        if (p.causal) cmax = min(cmax, qi + shift - k0);
                        cmax -= col_h;
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            if (2 * t > cmax) dk[t] = 0u;
                            else if (2 * t + 1 > cmax) dk[t] &= 0x0000ffffu;
                        }
                    }
                    FDBG(2);
                    if (k > 0) {   // this slot's previous product read dS: retired before the store
                        ptx::mbar_wait(&bar_acc[slot], (k - 1) & 1);
                        ptx::tc_fence_after_sync();
                    }
                    ptx::tmem_st_32x32b_x16(tDS, dk);
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&bar_p[slot]);
                    ++kown;
                    FDBG(3);
                }
                if (pending) {   // (also the flush pass after the last item, which has no sub-iteration)
                    read_out(rout_prev, qi_prev, n_prev);
                    pending = false;
                    FDBG(4);
                }
                v += 2;
            } while (v < v_end);
            if (!live) break;
            // item n is read out by group n & 1, after that group's first sub-iteration of the next item
            if ((n & 1) == grp) {
                pending = true;
                rout_prev = rout;
                qi_prev = qi;
                n_prev = n;
            }
            ++n;
        }
        FDBG_DUMP_ROW(my_iters, 0, 0);
        FDBG_DUMP_ROW(my_iters, 256, 1);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn2() {
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

// [B, T, W] bf16 view -> box 64 x 128 x 1, 128B swizzle, zero fill
int tmap_rows128(CUtensorMap* map, const void* base, int W, int T, int B, int rs, long long bs, int box_rows = 128) {
    EncodeTiledFn fn = encode_fn2();
    VLK_REQUIRE(fn != nullptr, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(rs) * 2, static_cast<cuuint64_t>(bs) * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VLK_REQUIRE(r == CUDA_SUCCESS, VLK_ERR_DRIVER, "cuTensorMapEncodeTiled(flash) failed with CUresult %d", (int)r);
    return VLK_OK;
}

constexpr int kFwdSmem = 6 * kTile + kFwdXchBytes + 128 + 1024;

}  // namespace

int attn_flash_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Tq, int Tk,
                   long long q_bs, int q_rs, long long k_bs, int k_rs, long long v_bs, int v_rs, long long o_bs,
                   int o_rs, int causal, float scale, cudaStream_t stream, int q_rows) {
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
        configured = true;
    }
    if (q_rows <= 0 || q_rows > Tq) q_rows = Tq;
    CUtensorMap tq, tk, tv;
    int rc = tmap_rows128(&tq, q, H * 64, Tq, B, q_rs, q_bs);
    if (rc) return rc;
    rc = tmap_rows128(&tk, k, H * 64, Tk, B, k_rs, k_bs);
    if (rc) return rc;
    rc = tmap_rows128(&tv, v, H * 64, Tk, B, v_rs, v_bs);
    if (rc) return rc;
    FlashFwdParams p;
    p.o = static_cast<bf16*>(o);
    p.lse = lse;
    p.os = Strides{o_bs, o_rs};
    p.B = B;
    p.H = H;
    p.Tq = Tq;
    p.Tk = Tk;
    p.causal = causal;
    p.q_rows = q_rows;
    p.nqb = (q_rows + BQ - 1) / BQ;
    p.wide_store = (reinterpret_cast<uintptr_t>(o) & 31u) == 0 && o_rs % 16 == 0 && o_bs % 16 == 0;
    p.div_hb = make_fastdiv(H * B);
    p.div_h = make_fastdiv(H);
    p.div_nqb = make_fastdiv(p.nqb);
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_attn_fwd: no sm_100 device");
    const long long items = static_cast<long long>(p.nqb) * H * B;
    const int grid = static_cast<int>(items < 2LL * sms ? items : 2LL * sms);   // persistent: two CTAs per SM
    p.div_grid = make_fastdiv(grid);
    flash_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, stream>>>(tq, tk, tv, p);
    VLK_CHECK_LAUNCH("vlk_attn_fwd(flash)");
    return VLK_OK;
}

int attn_flash_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                   void* dq, void* dk, void* dv, int B, int H, int Tq, int Tk, long long q_bs, int q_rs, long long k_bs,
                   int k_rs, long long v_bs, int v_rs, long long o_bs, int o_rs, long long dq_bs, int dq_rs,
                   long long dk_bs, int dk_rs, long long dv_bs, int dv_rs, int causal, float scale, float* delta,
                   cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(flash_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem));
        VLK_CUDA(cudaFuncSetAttribute(flash_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem));
        configured = true;
    }
    const int rows = B * H * Tq;
    flash_delta_kernel<<<(rows + 15) / 16, 128, 0, stream>>>(static_cast<const bf16*>(o), static_cast<const bf16*>(d_o),
                                                          delta, H, Tq, Strides{o_bs, o_rs}, rows);
    VLK_CHECK_LAUNCH("vlk_attn_bwd(delta)");
    CUtensorMap tq, tk, tv, tdo;
    int rc = tmap_rows128(&tq, q, H * 64, Tq, B, q_rs, q_bs);
    if (rc) return rc;
    rc = tmap_rows128(&tk, k, H * 64, Tk, B, k_rs, k_bs);
    if (rc) return rc;
    rc = tmap_rows128(&tv, v, H * 64, Tk, B, v_rs, v_bs);
    if (rc) return rc;
    rc = tmap_rows128(&tdo, d_o, H * 64, Tq, B, o_rs, o_bs);
    if (rc) return rc;
    FlashBwdParams p;
    p.lse = lse;
    p.delta = delta;
    p.H = H;
    p.Tq = Tq;
    p.Tk = Tk;
    p.causal = causal;
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.out0 = static_cast<bf16*>(dk);
    p.out1 = static_cast<bf16*>(dv);
    p.s0 = Strides{dk_bs, dk_rs};
    p.s1 = Strides{dv_bs, dv_rs};
    {
        CUtensorMap tq64, tdo64;   // 64-row boxes: the dK/dV kernel walks the queries in sub-blocks of 64
        rc = tmap_rows128(&tq64, q, H * 64, Tq, B, q_rs, q_bs, BQS);
        if (rc) return rc;
        rc = tmap_rows128(&tdo64, d_o, H * 64, Tq, B, o_rs, o_bs, BQS);
        if (rc) return rc;
        const int sms = device_sm_count();
        VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_attn_bwd: no sm_100 device");
        const long long items = static_cast<long long>((Tk + BK - 1) / BK) * H * B;
        const int grid = static_cast<int>(items < sms ? items : sms);   // persistent: one CTA per SM
        p.div_hb = make_fastdiv(H * B);
        p.div_h = make_fastdiv(H);
        p.div_grid = make_fastdiv(grid);
        flash_bwd_dkv_kernel<<<grid, kDkvThreads, kDkvSmem, stream>>>(tq64, tk, tv, tdo64, p, B);
    }
    VLK_CHECK_LAUNCH("vlk_attn_bwd(flash dkv)");
    p.out0 = static_cast<bf16*>(dq);
    p.out1 = nullptr;
    p.s0 = Strides{dq_bs, dq_rs};
    {
        CUtensorMap tk64, tv64;   // 64-row boxes: the dQ kernel walks the keys in sub-blocks of 64
        rc = tmap_rows128(&tk64, k, H * 64, Tk, B, k_rs, k_bs, BQS);
        if (rc) return rc;
        rc = tmap_rows128(&tv64, v, H * 64, Tk, B, v_rs, v_bs, BQS);
        if (rc) return rc;
        const int sms = device_sm_count();
        const long long items = static_cast<long long>((Tq + BQ - 1) / BQ) * H * B;
        const int grid = static_cast<int>(items < sms ? items : sms);   // persistent: one CTA per SM
        p.div_grid = make_fastdiv(grid);
        flash_bwd_dq_kernel<<<grid, kDkvThreads, kDqSmem, stream>>>(tq, tk64, tv64, tdo, p, B);
    }
    VLK_CHECK_LAUNCH("vlk_attn_bwd(flash dq)");
    return VLK_OK;
}

}  // namespace vlk

#ifdef VLK_BRINGUP
extern "C" int vlk_debug_flash_dump(long long* host_out, int n) {
    if (n > 64 * 16) n = 64 * 16;
    return static_cast<int>(cudaMemcpyFromSymbol(host_out, vlk::g_flash_dbg, sizeof(long long) * n));
}
#endif
