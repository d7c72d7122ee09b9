// Shared host/device helpers for libvlk: error convention, bf16 packing, warp reductions.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/vlk.h"

namespace vlk {

// Thread-local last-error text, returned by vlk_last_error_string().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
void count_launch();

#define VLK_REQUIRE(cond, code, ...)                     \
    do {                                                 \
        if (!(cond)) return ::vlk::set_error((code), __VA_ARGS__); \
    } while (0)

// Check the launch itself (not completion): kernels are asynchronous by contract.
#define VLK_CHECK_LAUNCH(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return ::vlk::set_error(static_cast<int>(e__), "%s: launch failed: %s", (name), \
                                    cudaGetErrorString(e__));                           \
        ::vlk::count_launch();                                                          \
    } while (0)

#define VLK_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return ::vlk::set_error(static_cast<int>(e__), "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int device_sm_count();

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 8 bf16 <-> 8 floats through one 16-byte vector.
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const bf162* h = reinterpret_cast<const bf162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    bf162* h = reinterpret_cast<bf162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// Counter-based RNG for dropout: Philox-4x32-10 (Salmon et al. 2011), the generator torch's CUDA dropout uses.
// Keyed by a device-resident (seed, step) pair so that a captured CUDA graph draws fresh masks on every replay,
// and by a per-call-site stream id so forward and backward of one site regenerate the SAME mask.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
struct DropoutKey {
    uint2 key;        // seed (low, high)
    uint32_t stream;  // call-site id
    uint32_t step;    // device step counter
    uint32_t thresh;  // drop if rand < thresh
    float inv_keep;   // 1 / (1 - p)
};
__device__ __forceinline__ DropoutKey make_dropout_key(const unsigned long long* seed_state, uint32_t stream, float p) {
    DropoutKey d;
    const unsigned long long seed = seed_state[0];
    d.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    d.stream = stream;
    d.step = static_cast<uint32_t>(seed_state[1]);
    d.thresh = static_cast<uint32_t>(fminf(p, 0.999999f) * 4294967296.0f);
    d.inv_keep = 1.0f / (1.0f - p);
    return d;
}
// four keep-scales (0 or 1/(1-p)) for elements 4*group .. 4*group+3
__device__ __forceinline__ void dropout_scales4(const DropoutKey& d, unsigned long long group, float (&m)[4]) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(group), static_cast<uint32_t>(group >> 32), d.stream,
                                             d.step), d.key);
    m[0] = r.x >= d.thresh ? d.inv_keep : 0.f;
    m[1] = r.y >= d.thresh ? d.inv_keep : 0.f;
    m[2] = r.z >= d.thresh ? d.inv_keep : 0.f;
    m[3] = r.w >= d.thresh ? d.inv_keep : 0.f;
}
// single element (small attention tiles): element index -> its scale
__device__ __forceinline__ float dropout_scale1(const DropoutKey& d, unsigned long long idx) {
    float m[4];
    dropout_scales4(d, idx >> 2, m);
    const int s = static_cast<int>(idx & 3);
    return s == 0 ? m[0] : (s == 1 ? m[1] : (s == 2 ? m[2] : m[3]));
}

// The three GELU flavours on the path: tanh (GPT-2), erf (Q-Former), quick (CLIP).
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float act_apply(int act, float x) {
    switch (act) {
        case VLK_ACT_GELU_TANH: {
            const float k0 = 0.7978845608028654f, k1 = 0.044715f;
            float u = k0 * (x + k1 * x * x * x);
            return 0.5f * x * (1.0f + tanh_fast(u));  // one MUFU op; 2^-11 relative error << bf16 rounding
        }
        case VLK_ACT_GELU_ERF:
            return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
        case VLK_ACT_QUICK_GELU:
            // x * sigmoid(1.702 x) with sigmoid(z) = 0.5 tanh(z/2) + 0.5: one MUFU op instead of ex2 + rcp
            return x * (0.5f * tanh_fast(0.851f * x) + 0.5f);
        default:
            return x;
    }
}
// d act(x) / dx
__device__ __forceinline__ float act_grad(int act, float x) {
    switch (act) {
        case VLK_ACT_GELU_TANH: {
            const float k0 = 0.7978845608028654f, k1 = 0.044715f;
            float u = k0 * (x + k1 * x * x * x);
            float t = tanh_fast(u);
            float du = k0 * (1.0f + 3.0f * k1 * x * x);
            return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
        }
        case VLK_ACT_GELU_ERF: {
            float cdf = 0.5f * (1.0f + erff(x * 0.7071067811865476f));
            float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
            return cdf + x * pdf;
        }
        case VLK_ACT_QUICK_GELU: {
            float s = 0.5f * tanh_fast(0.851f * x) + 0.5f;
            return s + 1.702f * x * s * (1.0f - s);
        }
        default:
            return 1.0f;
    }
}

}  // namespace vlk
