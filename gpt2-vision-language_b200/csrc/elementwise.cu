// Memory-bound kernels of the captioning / GPT-2 step: 257->33 pooling + L2 normalise, embedding gather +
// concat, CLIP patch unfold / sequence assembly, column sums, transposes, casts, gate gradient, argmax.
// All are HBM-bound: 16-byte vector accesses, coalesced along the feature dimension.
#include "common.cuh"

namespace vlk {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red /* >= 32 floats */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (warp == 0) v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
    __syncthreads();
    return red[0];
}

// ---------------------------------------------------------------------------------------------------
// 257 -> 33 pooling (+ L2 normalise).  One block per (image, output token); thread = one 8-wide chunk.
// Output token 0 = CLS; token 1 + 8r + c = mean of patch rows 4r..4r+3, cols 2c..2c+1 of the 16x16 grid.
// ---------------------------------------------------------------------------------------------------
template <bool FP32>
__global__ void pool33_kernel(const void* __restrict__ in_, void* __restrict__ out_, int D, int normalize) {
    __shared__ float red[32];
    const int b = blockIdx.y, j = blockIdx.x;
    const int col = threadIdx.x * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto load = [&](int token, float (&f)[8]) {
        const size_t off = (static_cast<size_t>(b) * 257 + token) * D + col;
        if (FP32) {
            const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(in_) + off);
            float4 a = __ldg(p), c = __ldg(p + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
            f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
        } else {
            unpack8(ldg16(static_cast<const bf16*>(in_) + off), f);
        }
    };
    if (col < D) {
        if (j == 0) {
            load(0, acc);
        } else {
            const int r = (j - 1) >> 3, c = (j - 1) & 7;
#pragma unroll
            for (int dr = 0; dr < 4; ++dr) {
#pragma unroll
                for (int dc = 0; dc < 2; ++dc) {
                    float f[8];
                    load(1 + 16 * (4 * r + dr) + (2 * c + dc), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += f[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= 0.125f;
        }
    }
    if (normalize) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += acc[i] * acc[i];
        ss = block_sum(ss, red);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= inv;
    }
    if (col < D) {
        const size_t off = (static_cast<size_t>(b) * 33 + j) * D + col;
        if (FP32) {
            float4* p = reinterpret_cast<float4*>(static_cast<float*>(out_) + off);
            p[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            p[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
            stg16(static_cast<bf16*>(out_) + off, pack8(acc));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Embedding gather + position add + image-prefix concat.  One block per output row.
// ---------------------------------------------------------------------------------------------------
__global__ void embed_concat_kernel(const long long* __restrict__ ids, const bf16* __restrict__ wte,
                                    const bf16* __restrict__ wpe, const bf16* __restrict__ prefix,
                                    bf16* __restrict__ out, int T, int prefix_len, int C, int pos0, int vocab) {
    const int L = prefix_len + T;
    const int b = blockIdx.x / L, t = blockIdx.x % L;
    bf16* o = out + static_cast<size_t>(blockIdx.x) * C;
    if (t < prefix_len) {
        const bf16* p = prefix + (static_cast<size_t>(b) * prefix_len + t) * C;
        for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) stg16(o + col, ldg16(p + col));
    } else {
        const int tt = t - prefix_len;
        const long long id = ids[static_cast<size_t>(b) * T + tt];
        if (id < 0 || id >= vocab) {
            // out-of-vocabulary id: torch's embedding raises a device assert; here the row is poisoned with NaN so the
            // loss of the step is NaN instead of a silent read outside wte
            const uint4 nan8 = make_uint4(0x7FC07FC0u, 0x7FC07FC0u, 0x7FC07FC0u, 0x7FC07FC0u);
            for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) stg16(o + col, nan8);
            return;
        }
        const bf16* e = wte + static_cast<size_t>(id) * C;
        const bf16* p = wpe + static_cast<size_t>(pos0 + tt) * C;
        for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) {
            float a[8], c[8];
            unpack8(ldg16(e + col), a);
            unpack8(ldg16(p + col), c);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += c[i];
            stg16(o + col, pack8(a));
        }
    }
}

// dwte[ids[b,t]] += dout[b, prefix+t];  dwpe[t] += dout[b, prefix+t]   (fp32 atomics; pretraining only)
__global__ void embed_bwd_kernel(const long long* __restrict__ ids, const bf16* __restrict__ dout,
                                 float* __restrict__ dwte, float* __restrict__ dwpe, int T, int prefix_len, int C) {
    const int b = blockIdx.x / T, t = blockIdx.x % T;
    const bf16* g = dout + (static_cast<size_t>(b) * (prefix_len + T) + prefix_len + t) * C;
    const long long id = ids[blockIdx.x];
    for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) {
        float f[8];
        unpack8(ldg16(g + col), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (dwte) atomicAdd(dwte + static_cast<size_t>(id) * C + col + i, f[i]);
            if (dwpe) atomicAdd(dwpe + static_cast<size_t>(t) * C + col + i, f[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Embedding gradient accumulated into bf16 rows without dense [V, C] temporaries (vlk_embed_bwd_acc).
//   A: first[id] = min position holding that id          B: scratch[first[id_p]] += dout[p]   (fp32 atomics)
//   C: positions that ARE a first position add their scratch row into dwte[id] (one owner per row: plain RMW),
//      then restore the workspaces (scratch row = 0, first[id] = INT_MAX)
//   D: dwpe[t] += sum_b dout[b, t]
// ---------------------------------------------------------------------------------------------------
__global__ void embed_first_pos_kernel(const long long* __restrict__ ids, int* __restrict__ first, int n, int vocab) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long id = ids[p];
    if (id >= 0 && id < vocab) atomicMin(first + id, p);
}
__global__ void embed_scatter_kernel(const long long* __restrict__ ids, const bf16* __restrict__ dout,
                                     const int* __restrict__ first, float* __restrict__ scratch, int T, int prefix_len,
                                     int C, int vocab) {
    const int p = blockIdx.x, b = p / T, t = p % T;
    const long long id = ids[p];
    if (id < 0 || id >= vocab) return;
    const bf16* g = dout + (static_cast<size_t>(b) * (prefix_len + T) + prefix_len + t) * C;
    float* dst = scratch + static_cast<size_t>(first[id]) * C;
    for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) {
        float f[8];
        unpack8(ldg16(g + col), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + col + i, f[i]);
    }
}
__global__ void embed_flush_kernel(const long long* __restrict__ ids, int* __restrict__ first,
                                   float* __restrict__ scratch, bf16* __restrict__ dwte, int C, int vocab) {
    const int p = blockIdx.x;
    const long long id = ids[p];
    if (id < 0 || id >= vocab || first[id] != p) return;     // block-uniform
    float* src = scratch + static_cast<size_t>(p) * C;
    bf16* dst = dwte + static_cast<size_t>(id) * C;
    for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(dst + col), a);
        const float4 s0 = *reinterpret_cast<const float4*>(src + col), s1 = *reinterpret_cast<const float4*>(src + col + 4);
        a[0] += s0.x; a[1] += s0.y; a[2] += s0.z; a[3] += s0.w;
        a[4] += s1.x; a[5] += s1.y; a[6] += s1.z; a[7] += s1.w;
        stg16(dst + col, pack8(a));
        *reinterpret_cast<float4*>(src + col) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(src + col + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (threadIdx.x == 0) first[id] = 0x7fffffff;
}
__global__ void embed_wpe_grad_kernel(const bf16* __restrict__ dout, bf16* __restrict__ dwpe, int B, int T,
                                      int prefix_len, int C) {
    const int t = blockIdx.x;
    for (int col = threadIdx.x * 8; col < C; col += blockDim.x * 8) {
        float acc[8];
        unpack8(*reinterpret_cast<const uint4*>(dwpe + static_cast<size_t>(t) * C + col), acc);
        for (int b = 0; b < B; ++b) {
            float f[8];
            unpack8(ldg16(dout + (static_cast<size_t>(b) * (prefix_len + T) + prefix_len + t) * C + col), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
        stg16(dwpe + static_cast<size_t>(t) * C + col, pack8(acc));
    }
}

// ---------------------------------------------------------------------------------------------------
// CLIP patch unfold: pixels [B,3,224,224] -> rows [B*256, Kpad], k = c*196 + i*14 + j.
// One block per (image, patch row); thread x walks one pixel column of the 224-wide stripe.
// ---------------------------------------------------------------------------------------------------
template <bool FP32>
__global__ void im2col_patch14_kernel(const void* __restrict__ pixels, bf16* __restrict__ out, int Kpad) {
    const int b = blockIdx.y, py = blockIdx.x;
    const int x = threadIdx.x;  // 0..255
    const size_t row0 = (static_cast<size_t>(b) * 256 + py * 16);
    if (x < 224) {
        const int px = x / 14, j = x % 14;
        bf16* o = out + (row0 + px) * Kpad;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                const size_t off = ((static_cast<size_t>(b) * 3 + c) * 224 + (py * 14 + i)) * 224 + x;
                float v = FP32 ? __ldg(static_cast<const float*>(pixels) + off)
                               : __bfloat162float(static_cast<const bf16*>(pixels)[off]);
                o[c * 196 + i * 14 + j] = __float2bfloat16(v);
            }
        }
    }
    // zero the K padding (588..Kpad) of the 16 rows
    const int pad = Kpad - 588;
    for (int idx = threadIdx.x; idx < 16 * pad; idx += blockDim.x)
        out[(row0 + idx / pad) * Kpad + 588 + idx % pad] = __float2bfloat16(0.f);
}

// x[b,0] = cls + pos[0]; x[b,1+p] = patch[b*256+p] + pos[1+p]
__global__ void clip_assemble_kernel(const bf16* __restrict__ patch, const bf16* __restrict__ cls,
                                     const bf16* __restrict__ pos, bf16* __restrict__ out, int D) {
    const int b = blockIdx.x / 257, t = blockIdx.x % 257;
    const bf16* src = (t == 0) ? cls : patch + (static_cast<size_t>(b) * 256 + (t - 1)) * D;
    const bf16* p = pos + static_cast<size_t>(t) * D;
    bf16* o = out + static_cast<size_t>(blockIdx.x) * D;
    for (int col = threadIdx.x * 8; col < D; col += blockDim.x * 8) {
        float a[8], c[8];
        unpack8(ldg16(src + col), a);
        unpack8(ldg16(p + col), c);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += c[i];
        stg16(o + col, pack8(a));
    }
}

// ---------------------------------------------------------------------------------------------------
// Column sums (bias gradients): block = 32 lanes x 32 row-groups (1024 threads), each lane owns 8 columns, eight
// 16-byte loads in flight per thread.  At most 32 blocks share a column (gridDim.y <= 32): every block ends with one
// atomicAdd per column, and with ~300 blocks per column that same-address tail, not the stream, set the time
// (26 us for 25 MB).
// ---------------------------------------------------------------------------------------------------
constexpr int kColsumGroups = 32;
// grad_out != nullptr: `out` is a zeroed fp32 workspace; the last of the gridDim.y blocks of a 256-column group (counters,
// one zeroed uint32 per group) ADDS the finished sums into the bf16 gradient and zeroes workspace and counter again.
__global__ void __launch_bounds__(1024) colsum_kernel(const bf16* __restrict__ X, float* __restrict__ out, int rows,
                                                      int cols, int ldx, bf16* __restrict__ grad_out,
                                                      unsigned int* __restrict__ counters) {
    __shared__ float red[kColsumGroups][32 * 8 + 1];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + lane) * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (col < cols) {
        const int stride = gridDim.y * kColsumGroups;
        int r = blockIdx.y * kColsumGroups + g;
        for (; r + 7 * stride < rows; r += 8 * stride) {
            uint4 q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = ldg16(X + static_cast<size_t>(r + u * stride) * ldx + col);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float f[8];
                unpack8(q[u], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += f[i];
            }
        }
        for (; r < rows; r += stride) {
            float f[8];
            unpack8(ldg16(X + static_cast<size_t>(r) * ldx + col), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[g][lane * 8 + i] = acc[i];
    __syncthreads();
    // 256 columns x 32 partials: thread t sums column t & 255 over a quarter of the groups, 4 atomics per column
    {
        const int c = threadIdx.x & 255, part = threadIdx.x >> 8;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kColsumGroups / 4; ++w) s += red[part * (kColsumGroups / 4) + w][c];
        const int gc = blockIdx.x * 256 + c;
        if (gc < cols) atomicAdd(out + gc, s);
    }
    if (grad_out != nullptr) {
        __shared__ int is_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(counters + blockIdx.x, 1u) == gridDim.y - 1;
        __syncthreads();
        if (is_last) {
            __threadfence();
            const int gc = blockIdx.x * 256 + threadIdx.x;
            if (threadIdx.x < 256 && gc < cols) {
                const float v = __ldcg(out + gc);
                out[gc] = 0.f;
                grad_out[gc] = __float2bfloat16(v + __bfloat162float(grad_out[gc]));
            }
            if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
        }
    }
}

__global__ void transpose_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int rows, int cols,
                                 int ld_src, int ld_dst) {
    __shared__ bf16 tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = src[static_cast<size_t>(r0 + i) * ld_src + c];
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (c0 + i < cols && r < rows) dst[static_cast<size_t>(c0 + i) * ld_dst + r] = tile[threadIdx.x][i];
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y,
                           long long n8) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float x[8], z[8];
        unpack8(ldg16(a + i * 8), x);
        unpack8(ldg16(b + i * 8), z);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] += z[k];
        stg16(y + i * 8, pack8(x));
    }
}

__global__ void cast_f2b_kernel(const float* __restrict__ s, bf16* __restrict__ d, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        d[i] = __float2bfloat16(s[i]);
}
__global__ void cast_b2f_kernel(const bf16* __restrict__ s, float* __restrict__ d, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        d[i] = __bfloat162float(s[i]);
}

// out[0] += (1 - tanh(gate)^2) * sum(dy * y)
__global__ void gate_grad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                 const float* __restrict__ gate, float* __restrict__ out, long long n8) {
    __shared__ float red[32];
    float acc = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float a[8], b[8];
        unpack8(ldg16(dy + i * 8), a);
        unpack8(ldg16(y + i * 8), b);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += a[k] * b[k];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const float t = tanhf(*gate);
        atomicAdd(out, acc * (1.0f - t * t));
    }
}

// first index of the row maximum (matches torch.argmax tie-breaking on the lowest index)
__global__ void __launch_bounds__(256) argmax_kernel(const bf16* __restrict__ logits, long long* __restrict__ out,
                                                     int V, int ld) {
    __shared__ float sv[8];
    __shared__ int si[8];
    const bf16* r = logits + static_cast<size_t>(blockIdx.x) * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = threadIdx.x * 8; c < V; c += blockDim.x * 8) {
        float f[8];
        unpack8(ldg16(r + c), f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c + i < V && (f[i] > best || (f[i] == best && c + i < bi))) {
                best = f[i];
                bi = c + i;
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) {
            best = ov;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = best;
        si[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) {
                best = sv[w];
                bi = si[w];
            }
        out[blockIdx.x] = bi;
    }
}

// y = residual + x * keep / (1 - p)   (residual may be NULL); 8 elements per thread, two Philox draws
__global__ void dropout_add_kernel(const bf16* __restrict__ x, const bf16* __restrict__ residual,
                                   bf16* __restrict__ y, long long n8, float p,
                                   const unsigned long long* __restrict__ seed_state, uint32_t stream) {
    const DropoutKey dk = make_dropout_key(seed_state, stream, p);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float a[8], m0[4], m1[4];
        unpack8(ldg16(x + i * 8), a);
        dropout_scales4(dk, static_cast<unsigned long long>(i) * 2, m0);
        dropout_scales4(dk, static_cast<unsigned long long>(i) * 2 + 1, m1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] *= m0[k];
            a[4 + k] *= m1[k];
        }
        if (residual != nullptr) {
            float r[8];
            unpack8(ldg16(residual + i * 8), r);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] += r[k];
        }
        stg16(y + i * 8, pack8(a));
    }
}

inline int grid_for(long long work_items, int threads, int sms) {
    long long b = (work_items + threads - 1) / threads;
    long long cap = static_cast<long long>(sms) * 16;
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------------
// A scalar riding in the gradient all-reduce: Q8.24 fixed point, 8 base-16 digits, one per bucket slot.
// Sums of <= 16 digits (< 256) and a division by a power of two are exact in bf16, so the averaged value is
// recovered exactly (to 2^-24) whatever the reduction order of the collective.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void scalar_pack_digits_kernel(const float* __restrict__ v, T* __restrict__ slots) {
    const float x = fminf(fmaxf(*v, 0.f), 255.99999f);
    const unsigned int q = static_cast<unsigned int>(llrintf(x * 16777216.0f));
    const float d = static_cast<float>((q >> (4 * threadIdx.x)) & 15u);
    slots[threadIdx.x] = static_cast<T>(d);
    if (threadIdx.x == 0 && !(*v == *v)) slots[7] = static_cast<T>(nanf(""));   // a NaN loss stays a NaN
}
template <typename T>
__global__ void scalar_unpack_digits_kernel(const T* __restrict__ slots, float* __restrict__ v) {
    double acc = 0.0, w = 1.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc += static_cast<double>(static_cast<float>(slots[i])) * w;
        w *= 16.0;
    }
    *v = static_cast<float>(acc * (1.0 / 16777216.0));
}

}  // namespace
}  // namespace vlk

using namespace vlk;

extern "C" int vlk_pool33_l2norm(const void* in, void* out, int B, int D, int in_fp32, int normalize, void* stream) {
    VLK_REQUIRE(in && out && B > 0, VLK_ERR_INVALID_ARG, "vlk_pool33_l2norm: null pointer or B=%d", B);
    VLK_REQUIRE(D > 0 && D % 8 == 0 && D <= 8192, VLK_ERR_INVALID_ARG, "vlk_pool33_l2norm: D=%d", D);
    VLK_REQUIRE(aligned16(in) && aligned16(out), VLK_ERR_ALIGNMENT, "vlk_pool33_l2norm: 16B alignment");
    const int threads = ((D / 8 + 31) / 32) * 32;
    const dim3 grid(33, B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (in_fp32)
        pool33_kernel<true><<<grid, threads, 0, s>>>(in, out, D, normalize);
    else
        pool33_kernel<false><<<grid, threads, 0, s>>>(in, out, D, normalize);
    VLK_CHECK_LAUNCH("vlk_pool33_l2norm");
    return VLK_OK;
}

extern "C" int vlk_embed_concat_fwd(const long long* ids, const void* wte, const void* wpe, const void* prefix,
                                    void* out, int B, int T, int prefix_len, int C, int pos0, int vocab, void* stream) {
    VLK_REQUIRE(ids && wte && wpe && out, VLK_ERR_INVALID_ARG, "vlk_embed_concat_fwd: null pointer");
    VLK_REQUIRE(B > 0 && T > 0 && prefix_len >= 0 && C % 8 == 0 && vocab > 0, VLK_ERR_INVALID_ARG,
                "vlk_embed_concat_fwd: B=%d T=%d prefix=%d C=%d vocab=%d", B, T, prefix_len, C, vocab);
    VLK_REQUIRE(prefix_len == 0 || prefix, VLK_ERR_INVALID_ARG, "vlk_embed_concat_fwd: prefix missing");
    const int threads = C / 8 >= 128 ? 128 : ((C / 8 + 31) / 32) * 32;
    embed_concat_kernel<<<B * (prefix_len + T), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, static_cast<const bf16*>(wte), static_cast<const bf16*>(wpe), static_cast<const bf16*>(prefix),
        static_cast<bf16*>(out), T, prefix_len, C, pos0, vocab);
    VLK_CHECK_LAUNCH("vlk_embed_concat_fwd");
    return VLK_OK;
}

extern "C" int vlk_embed_bwd(const long long* ids, const void* dout, float* dwte, float* dwpe, int B, int T,
                             int prefix_len, int C, void* stream) {
    VLK_REQUIRE(ids && dout && (dwte || dwpe), VLK_ERR_INVALID_ARG, "vlk_embed_bwd: null pointer");
    VLK_REQUIRE(B > 0 && T > 0 && C % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_embed_bwd: shape");
    embed_bwd_kernel<<<B * T, 96, 0, static_cast<cudaStream_t>(stream)>>>(ids, static_cast<const bf16*>(dout), dwte,
                                                                         dwpe, T, prefix_len, C);
    VLK_CHECK_LAUNCH("vlk_embed_bwd");
    return VLK_OK;
}

extern "C" int vlk_embed_bwd_acc(const long long* ids, const void* dout, void* dwte, void* dwpe, int* first_pos,
                                 float* scratch, int B, int T, int prefix_len, int C, int vocab, void* stream) {
    VLK_REQUIRE(ids && dout && (dwte || dwpe), VLK_ERR_INVALID_ARG, "vlk_embed_bwd_acc: null pointer");
    VLK_REQUIRE(!dwte || (first_pos && scratch), VLK_ERR_INVALID_ARG, "vlk_embed_bwd_acc: workspaces missing");
    VLK_REQUIRE(B > 0 && T > 0 && C % 8 == 0 && vocab > 0, VLK_ERR_INVALID_ARG, "vlk_embed_bwd_acc: shape");
    VLK_REQUIRE(aligned16(dout) && (!dwte || aligned16(dwte)) && (!dwpe || aligned16(dwpe)) && (!scratch || aligned16(scratch)),
                VLK_ERR_ALIGNMENT, "vlk_embed_bwd_acc: 16B alignment");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = B * T;
    const int threads = C / 8 >= 96 ? 96 : ((C / 8 + 31) / 32) * 32;
    if (dwte) {
        embed_first_pos_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, first_pos, n, vocab);
        VLK_CHECK_LAUNCH("vlk_embed_bwd_acc(first)");
        embed_scatter_kernel<<<n, threads, 0, s>>>(ids, static_cast<const bf16*>(dout), first_pos, scratch, T, prefix_len, C,
                                                  vocab);
        VLK_CHECK_LAUNCH("vlk_embed_bwd_acc(scatter)");
        embed_flush_kernel<<<n, threads, 0, s>>>(ids, first_pos, scratch, static_cast<bf16*>(dwte), C, vocab);
        VLK_CHECK_LAUNCH("vlk_embed_bwd_acc(flush)");
    }
    if (dwpe) {
        embed_wpe_grad_kernel<<<T, threads, 0, s>>>(static_cast<const bf16*>(dout), static_cast<bf16*>(dwpe), B, T, prefix_len,
                                                   C);
        VLK_CHECK_LAUNCH("vlk_embed_bwd_acc(wpe)");
    }
    return VLK_OK;
}

extern "C" int vlk_im2col_patch14(const void* pixels, void* out, int B, int Kpad, int in_fp32, void* stream) {
    VLK_REQUIRE(pixels && out && B > 0, VLK_ERR_INVALID_ARG, "vlk_im2col_patch14: null pointer");
    VLK_REQUIRE(Kpad >= 588 && Kpad % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_im2col_patch14: Kpad=%d", Kpad);
    const dim3 grid(16, B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (in_fp32)
        im2col_patch14_kernel<true><<<grid, 256, 0, s>>>(pixels, static_cast<bf16*>(out), Kpad);
    else
        im2col_patch14_kernel<false><<<grid, 256, 0, s>>>(pixels, static_cast<bf16*>(out), Kpad);
    VLK_CHECK_LAUNCH("vlk_im2col_patch14");
    return VLK_OK;
}

extern "C" int vlk_clip_assemble(const void* patch, const void* cls, const void* pos, void* out, int B, int D,
                                 void* stream) {
    VLK_REQUIRE(patch && cls && pos && out && B > 0 && D % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_clip_assemble: args");
    clip_assemble_kernel<<<B * 257, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(patch), static_cast<const bf16*>(cls), static_cast<const bf16*>(pos),
        static_cast<bf16*>(out), D);
    VLK_CHECK_LAUNCH("vlk_clip_assemble");
    return VLK_OK;
}

extern "C" int vlk_colsum_bf16(const void* X, float* out, int rows, int cols, int ldx, void* stream) {
    VLK_REQUIRE(X && out && rows > 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0, VLK_ERR_INVALID_ARG,
                "vlk_colsum_bf16: rows=%d cols=%d ldx=%d", rows, cols, ldx);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VLK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
    const int gx = (cols + 255) / 256;
    int gy = (rows + 255) / 256;     // at least 8 rows per thread ...
    if (gy > 32) gy = 32;            // ... and at most 32 (x4) same-address atomics per column
    if (gy < 1) gy = 1;
    colsum_kernel<<<dim3(gx, gy), 1024, 0, s>>>(static_cast<const bf16*>(X), out, rows, cols, ldx, nullptr, nullptr);
    VLK_CHECK_LAUNCH("vlk_colsum_bf16");
    return VLK_OK;
}

extern "C" int vlk_colsum_bf16_acc(const void* X, float* workspace, unsigned int* counters, void* grad, int rows, int cols,
                                   int ldx, void* stream) {
    VLK_REQUIRE(X && workspace && counters && grad && rows > 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0,
                VLK_ERR_INVALID_ARG, "vlk_colsum_bf16_acc: rows=%d cols=%d ldx=%d", rows, cols, ldx);
    const int gx = (cols + 255) / 256;
    int gy = (rows + 255) / 256;
    if (gy > 32) gy = 32;
    if (gy < 1) gy = 1;
    colsum_kernel<<<dim3(gx, gy), 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(X), workspace, rows, cols, ldx, static_cast<bf16*>(grad), counters);
    VLK_CHECK_LAUNCH("vlk_colsum_bf16_acc");
    return VLK_OK;
}

extern "C" int vlk_transpose_bf16(const void* src, void* dst, int rows, int cols, int ld_src, int ld_dst,
                                  void* stream) {
    VLK_REQUIRE(src && dst && rows > 0 && cols > 0, VLK_ERR_INVALID_ARG, "vlk_transpose_bf16: args");
    transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(src), static_cast<bf16*>(dst), rows, cols, ld_src, ld_dst);
    VLK_CHECK_LAUNCH("vlk_transpose_bf16");
    return VLK_OK;
}

extern "C" int vlk_add_bf16(const void* a, const void* b, void* y, long long n, void* stream) {
    VLK_REQUIRE(a && b && y && n > 0 && n % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_add_bf16: n=%lld", n);
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_add_bf16: no sm_100 device");
    add_kernel<<<grid_for(n / 8, 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(y), n / 8);
    VLK_CHECK_LAUNCH("vlk_add_bf16");
    return VLK_OK;
}

extern "C" int vlk_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
    VLK_REQUIRE(src && dst && n > 0, VLK_ERR_INVALID_ARG, "vlk_cast_f32_to_bf16: args");
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_cast_f32_to_bf16: no sm_100 device");
    cast_f2b_kernel<<<grid_for(n, 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<bf16*>(dst), n);
    VLK_CHECK_LAUNCH("vlk_cast_f32_to_bf16");
    return VLK_OK;
}
extern "C" int vlk_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream) {
    VLK_REQUIRE(src && dst && n > 0, VLK_ERR_INVALID_ARG, "vlk_cast_bf16_to_f32: args");
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_cast_bf16_to_f32: no sm_100 device");
    cast_b2f_kernel<<<grid_for(n, 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(src), dst, n);
    VLK_CHECK_LAUNCH("vlk_cast_bf16_to_f32");
    return VLK_OK;
}

extern "C" int vlk_gate_grad(const void* dy, const void* y, const float* gate, float* out, long long n,
                             void* stream) {
    VLK_REQUIRE(dy && y && gate && out && n > 0 && n % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_gate_grad: n=%lld", n);
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_gate_grad: no sm_100 device");
    int grid = grid_for(n / 8, 256, sms);
    if (grid > sms * 2) grid = sms * 2;
    gate_grad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(dy),
                                                                        static_cast<const bf16*>(y), gate, out, n / 8);
    VLK_CHECK_LAUNCH("vlk_gate_grad");
    return VLK_OK;
}

extern "C" int vlk_argmax_rows(const void* logits, long long* out, int rows, int V, int ld, void* stream) {
    VLK_REQUIRE(logits && out && rows > 0 && V > 0 && V % 8 == 0 && ld % 8 == 0, VLK_ERR_INVALID_ARG,
                "vlk_argmax_rows: rows=%d V=%d ld=%d", rows, V, ld);
    argmax_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(logits), out, V, ld);
    VLK_CHECK_LAUNCH("vlk_argmax_rows");
    return VLK_OK;
}

extern "C" int vlk_dropout_add_bf16(const void* x, const void* residual, void* y, long long n, float p,
                                    const unsigned long long* seed_state, unsigned int stream_id, void* stream) {
    VLK_REQUIRE(x && y && seed_state && n > 0 && n % 8 == 0, VLK_ERR_INVALID_ARG, "vlk_dropout_add_bf16: n=%lld", n);
    VLK_REQUIRE(p >= 0.f && p < 1.f, VLK_ERR_INVALID_ARG, "vlk_dropout_add_bf16: p=%f", p);
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_dropout_add_bf16: no sm_100 device");
    dropout_add_kernel<<<grid_for(n / 8, 256, sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), static_cast<const bf16*>(residual), static_cast<bf16*>(y), n / 8, p, seed_state,
        stream_id);
    VLK_CHECK_LAUNCH("vlk_dropout_add_bf16");
    return VLK_OK;
}

extern "C" int vlk_scalar_pack_digits(const float* value, void* slots, int slots_fp32, void* stream) {
    VLK_REQUIRE(value && slots, VLK_ERR_INVALID_ARG, "vlk_scalar_pack_digits: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (slots_fp32)
        scalar_pack_digits_kernel<float><<<1, 8, 0, s>>>(value, static_cast<float*>(slots));
    else
        scalar_pack_digits_kernel<bf16><<<1, 8, 0, s>>>(value, static_cast<bf16*>(slots));
    VLK_CHECK_LAUNCH("vlk_scalar_pack_digits");
    return VLK_OK;
}

extern "C" int vlk_scalar_unpack_digits(const void* slots, float* value, int slots_fp32, void* stream) {
    VLK_REQUIRE(value && slots, VLK_ERR_INVALID_ARG, "vlk_scalar_unpack_digits: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (slots_fp32)
        scalar_unpack_digits_kernel<float><<<1, 1, 0, s>>>(static_cast<const float*>(slots), value);
    else
        scalar_unpack_digits_kernel<bf16><<<1, 1, 0, s>>>(static_cast<const bf16*>(slots), value);
    VLK_CHECK_LAUNCH("vlk_scalar_unpack_digits");
    return VLK_OK;
}
