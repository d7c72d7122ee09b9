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
