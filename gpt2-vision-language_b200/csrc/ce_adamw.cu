// Softmax cross-entropy over L2-resident logit chunks, and the fused clip-norm + AdamW update.
#include "common.cuh"

namespace vlk {
namespace {

constexpr int kCeThreads = 512;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float r = (lane < nw) ? red[lane] : (is_max ? -INFINITY : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    return r;  // every warp computes the same value
}

// One block per row.  Pass 1 streams the row (16 B vectors) into shared memory while tracking the max;
// pass 2 (from smem) accumulates sum(exp); pass 3 (from smem) writes (p - onehot) * w in place.
// The row is read from global exactly once and written at most once.
__global__ void __launch_bounds__(kCeThreads)
softmax_ce_rows_kernel(bf16* __restrict__ logits, const long long* __restrict__ labels,
                       const float* __restrict__ row_weight, float* __restrict__ loss_row,
                       const float* __restrict__ inv_count, int V, int ld, int write_grad) {
    extern __shared__ uint4 srow[];  // V/8 vectors
    __shared__ float red[32];
    const int row = blockIdx.x;
    bf16* r = logits + static_cast<size_t>(row) * ld;
    const long long label = labels[row];
    const int nvec = V / 8;
    if (label == -100 || label < 0 || label >= V) {
        // ignore_index (-100, F.cross_entropy's default and the value the reference masks with): zero loss, zero
        // gradient.  Any other label outside [0, V) is a caller bug (torch raises a device assert): the row's loss is
        // NaN, so the step's loss is NaN instead of a silent read past the cached row.
        if (threadIdx.x == 0) loss_row[row] = (label == -100) ? 0.f : nanf("");
        if (write_grad) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int i = threadIdx.x; i < nvec; i += blockDim.x) stg16(r + i * 8, z);
        }
        return;
    }
    float m = -INFINITY;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        const uint4 u = *reinterpret_cast<const uint4*>(r + i * 8);
        srow[i] = u;
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) m = fmaxf(m, f[k]);
    }
    m = block_reduce(m, red, true);
    float s = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[8];
        unpack8(srow[i], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += __expf(f[k] - m);
    }
    s = block_reduce(s, red, false);
    const float lse = m + __logf(s);
    if (threadIdx.x == 0) {
        const float xl = __bfloat162float(reinterpret_cast<const bf16*>(srow)[label]);
        loss_row[row] = lse - xl;
    }
    if (write_grad) {
        const float w = (row_weight ? row_weight[row] : 1.0f) * __ldg(inv_count);
        const int lvec = static_cast<int>(label >> 3), lsub = static_cast<int>(label & 7);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            float f[8];
            unpack8(srow[i], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __expf(f[k] - lse);
            if (i == lvec) f[lsub] -= 1.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] *= w;
            stg16(r + i * 8, pack8(f));
        }
    }
}

// out[1] = 1 / max(count, 1);   count = #(label >= 0)  or  sum(row_weight)
__global__ void ce_count_kernel(const long long* __restrict__ labels, const float* __restrict__ row_weight,
                                float* __restrict__ out, int rows) {
    __shared__ float red[32];
    float c = 0.f;
    for (int i = threadIdx.x; i < rows; i += blockDim.x)
        c += row_weight ? row_weight[i] : (labels[i] != -100 ? 1.f : 0.f);
    c = block_reduce(c, red, false);
    if (threadIdx.x == 0) out[1] = 1.0f / fmaxf(c, 1.0f);
}
// out[0] = sum(loss_row * w) * out[1]
__global__ void ce_finalize_kernel(const float* __restrict__ loss_row, const float* __restrict__ row_weight,
                                   float* __restrict__ out, int rows) {
    __shared__ float red[32];
    float c = 0.f;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) c += loss_row[i] * (row_weight ? row_weight[i] : 1.f);
    c = block_reduce(c, red, false);
    if (threadIdx.x == 0) out[0] = c * out[1];
}

// ---------------------------------------------------------------------------------------------------
// clip-norm + AdamW.  Table-driven multi-tensor kernels: blockIdx.y = tensor, blockIdx.x strides the tensor.
// 14 B per parameter of HBM traffic for bf16 params/grads/moments (p,g,m,v read; p,m,v written).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld1(const T* p, long long i);
template <>
__device__ __forceinline__ float ld1<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }
template <>
__device__ __forceinline__ float ld1<float>(const float* p, long long i) { return p[i]; }
__device__ __forceinline__ void st1(bf16* p, long long i, float v) { p[i] = __float2bfloat16(v); }
__device__ __forceinline__ void st1(float* p, long long i, float v) { p[i] = v; }

// Deterministic: every block writes its partial sum to partials[block]; grad_sumsq_finish_kernel adds them in a fixed
// order.  (With fp32 atomics the ranks of a data-parallel job computed norms that differed in the last bit, their clip
// factors differed, and the replicas' parameters drifted apart by bf16 ulps — tests/test_dp_gpu2.py.)
template <typename T>
__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const vlk_tensor_desc* __restrict__ table, float* __restrict__ partials) {
    __shared__ float red[32];
    const vlk_tensor_desc d = table[blockIdx.y];
    const T* g = static_cast<const T*>(d.grad);
    float acc = 0.f;
    if (g != nullptr) {
        constexpr int VEC = 16 / sizeof(T);
        // unaligned views fall back to the scalar tail loop
        const long long nvec = (reinterpret_cast<uintptr_t>(g) & 15u) ? 0 : d.numel / VEC;
        for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const uint4 u = ldg16(g + i * VEC);
            if (sizeof(T) == 2) {
                float f[8];
                unpack8(u, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += f[k] * f[k];
            } else {
                const float* f = reinterpret_cast<const float*>(&u);
                acc += f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
            }
        }
        if (blockIdx.x == 0)
            for (long long i = nvec * VEC + threadIdx.x; i < d.numel; i += blockDim.x) {
                const float v = ld1<T>(g, i);
                acc += v * v;
            }
    }
    acc = block_reduce(acc, red, false);
    if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// norm_sq[0] += sum of the n partials, in a fixed order: thread t adds partials t, t + 256, ... then a fixed tree.
__global__ void __launch_bounds__(256) grad_sumsq_finish_kernel(const float* __restrict__ partials, int n,
                                                                float* __restrict__ norm_sq) {
    __shared__ float sh[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) a += partials[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) norm_sq[0] += sh[0];
}

template <typename T>
__global__ void __launch_bounds__(256)
adamw_kernel(const vlk_tensor_desc* __restrict__ table, const float* __restrict__ norm_sq, float max_norm,
             const float* __restrict__ lr_p, float beta1, float beta2, float eps, const float* __restrict__ step_p) {
    const vlk_tensor_desc d = table[blockIdx.y];
    if (d.grad == nullptr) return;
    T* p = static_cast<T*>(d.param);
    const T* g = static_cast<const T*>(d.grad);
    T* m = static_cast<T*>(d.exp_avg);
    T* v = static_cast<T*>(d.exp_avg_sq);
    const float lr = __ldg(lr_p), step = __ldg(step_p);
    float clip = 1.0f;
    if (max_norm > 0.f && norm_sq != nullptr) {
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
        clip = fminf(1.0f, max_norm / (sqrtf(__ldg(norm_sq)) + 1e-6f));
    }
    const float bc1 = 1.0f - powf(beta1, step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, step));
    const float step_size = lr / bc1;
    const float decay = 1.0f - lr * d.weight_decay;
    auto update = [&](float gi, float& pi, float& mi, float& vi) {
        gi *= clip;
        pi *= decay;
        mi = beta1 * mi + (1.0f - beta1) * gi;
        vi = beta2 * vi + (1.0f - beta2) * gi * gi;
        pi -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    };
    // 16-byte accesses on all four streams (14 B/param of HBM traffic in bf16); unaligned views and the tail of a
    // tensor whose size is not a multiple of the vector width take the scalar loop
    constexpr int VEC = 16 / sizeof(T);
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                           reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
    const long long nvec = aligned ? d.numel / VEC : 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 ug = ldg16(g + i * VEC);
        const uint4 up = *reinterpret_cast<const uint4*>(p + i * VEC);
        const uint4 um = *reinterpret_cast<const uint4*>(m + i * VEC);
        const uint4 uv = *reinterpret_cast<const uint4*>(v + i * VEC);
        if constexpr (sizeof(T) == 2) {
            float fg[8], fp[8], fm[8], fv[8];
            unpack8(ug, fg);
            unpack8(up, fp);
            unpack8(um, fm);
            unpack8(uv, fv);
#pragma unroll
            for (int k = 0; k < 8; ++k) update(fg[k], fp[k], fm[k], fv[k]);
            *reinterpret_cast<uint4*>(p + i * VEC) = pack8(fp);
            *reinterpret_cast<uint4*>(m + i * VEC) = pack8(fm);
            *reinterpret_cast<uint4*>(v + i * VEC) = pack8(fv);
        } else {
            float4 fg = *reinterpret_cast<const float4*>(&ug), fp = *reinterpret_cast<const float4*>(&up);
            float4 fm = *reinterpret_cast<const float4*>(&um), fv = *reinterpret_cast<const float4*>(&uv);
            update(fg.x, fp.x, fm.x, fv.x);
            update(fg.y, fp.y, fm.y, fv.y);
            update(fg.z, fp.z, fm.z, fv.z);
            update(fg.w, fp.w, fm.w, fv.w);
            *reinterpret_cast<float4*>(p + i * VEC) = fp;
            *reinterpret_cast<float4*>(m + i * VEC) = fm;
            *reinterpret_cast<float4*>(v + i * VEC) = fv;
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * VEC + threadIdx.x; i < d.numel; i += blockDim.x) {
            float pi = ld1<T>(p, i), mi = ld1<T>(m, i), vi = ld1<T>(v, i);
            update(ld1<T>(g, i), pi, mi, vi);
            st1(p, i, pi);
            st1(m, i, mi);
            st1(v, i, vi);
        }
    }
}

}  // namespace
}  // namespace vlk

using namespace vlk;

extern "C" int vlk_softmax_ce_rows(void* logits, const long long* labels, const float* row_weight, float* loss_row,
                                   const float* inv_count, int rows, int V, int ld, int write_grad, void* stream) {
    VLK_REQUIRE(logits && labels && loss_row, VLK_ERR_INVALID_ARG, "vlk_softmax_ce_rows: null pointer");
    VLK_REQUIRE(!write_grad || inv_count, VLK_ERR_INVALID_ARG, "vlk_softmax_ce_rows: write_grad needs inv_count");
    VLK_REQUIRE(rows > 0 && V > 0 && V % 8 == 0 && ld % 8 == 0 && ld >= V, VLK_ERR_INVALID_ARG,
                "vlk_softmax_ce_rows: rows=%d V=%d ld=%d", rows, V, ld);
    VLK_REQUIRE(aligned16(logits), VLK_ERR_ALIGNMENT, "vlk_softmax_ce_rows: 16B alignment");
    const size_t smem = static_cast<size_t>(V) * 2;
    VLK_REQUIRE(smem <= 200 * 1024, VLK_ERR_UNSUPPORTED, "vlk_softmax_ce_rows: V=%d too large for the smem row cache", V);
    static bool configured = false;
    if (!configured) {
        VLK_CUDA(cudaFuncSetAttribute(softmax_ce_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    softmax_ce_rows_kernel<<<rows, kCeThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<bf16*>(logits), labels, row_weight, loss_row, inv_count, V, ld, write_grad);
    VLK_CHECK_LAUNCH("vlk_softmax_ce_rows");
    return VLK_OK;
}

extern "C" int vlk_ce_count(const long long* labels, const float* row_weight, float* out, int rows, void* stream) {
    VLK_REQUIRE((labels || row_weight) && out && rows > 0, VLK_ERR_INVALID_ARG, "vlk_ce_count: args");
    ce_count_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(labels, row_weight, out, rows);
    VLK_CHECK_LAUNCH("vlk_ce_count");
    return VLK_OK;
}

extern "C" int vlk_ce_finalize(const float* loss_row, const float* row_weight, float* out, int rows, void* stream) {
    VLK_REQUIRE(loss_row && out && rows > 0, VLK_ERR_INVALID_ARG, "vlk_ce_finalize: args");
    ce_finalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(loss_row, row_weight, out, rows);
    VLK_CHECK_LAUNCH("vlk_ce_finalize");
    return VLK_OK;
}

static int optimizer_grid_x(long long max_numel, int sms, int n_tensors) {
    long long bx = (max_numel + 256 * 8 - 1) / (256 * 8);
    long long cap = (static_cast<long long>(sms) * 8 + n_tensors - 1) / n_tensors;
    if (cap < 1) cap = 1;
    if (bx > cap) bx = cap;
    return static_cast<int>(bx < 1 ? 1 : bx);
}

extern "C" long long vlk_grad_sumsq_workspace_floats(int n_tensors, long long max_numel) {
    if (n_tensors <= 0 || max_numel <= 0) return -1;
    int sms = device_sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<long long>(optimizer_grid_x(max_numel, sms, n_tensors)) * n_tensors;
}

extern "C" int vlk_grad_sumsq(const vlk_tensor_desc* table, int n_tensors, long long max_numel, int dtype_fp32,
                              float* norm_sq, float* partials, void* stream) {
    VLK_REQUIRE(table && norm_sq && partials && n_tensors > 0 && n_tensors <= 65535 && max_numel > 0, VLK_ERR_INVALID_ARG,
                "vlk_grad_sumsq: n_tensors=%d", n_tensors);
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_grad_sumsq: no sm_100 device");
    const dim3 grid(optimizer_grid_x(max_numel, sms, n_tensors), n_tensors);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype_fp32)
        grad_sumsq_kernel<float><<<grid, 256, 0, s>>>(table, partials);
    else
        grad_sumsq_kernel<bf16><<<grid, 256, 0, s>>>(table, partials);
    VLK_CHECK_LAUNCH("vlk_grad_sumsq");
    grad_sumsq_finish_kernel<<<1, 256, 0, s>>>(partials, static_cast<int>(grid.x * grid.y), norm_sq);
    VLK_CHECK_LAUNCH("vlk_grad_sumsq(finish)");
    return VLK_OK;
}

extern "C" int vlk_adamw_step(const vlk_tensor_desc* table, int n_tensors, long long max_numel, int dtype_fp32,
                              const float* norm_sq, float max_norm, const float* lr, float beta1, float beta2,
                              float eps, const float* step, void* stream) {
    VLK_REQUIRE(table && lr && step && n_tensors > 0 && n_tensors <= 65535 && max_numel > 0, VLK_ERR_INVALID_ARG,
                "vlk_adamw_step: n_tensors=%d", n_tensors);
    VLK_REQUIRE(max_norm <= 0.f || norm_sq, VLK_ERR_INVALID_ARG, "vlk_adamw_step: clipping needs norm_sq");
    const int sms = device_sm_count();
    VLK_REQUIRE(sms > 0, VLK_ERR_ARCH, "vlk_adamw_step: no sm_100 device");
    const dim3 grid(optimizer_grid_x(max_numel, sms, n_tensors), n_tensors);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype_fp32)
        adamw_kernel<float><<<grid, 256, 0, s>>>(table, norm_sq, max_norm, lr, beta1, beta2, eps, step);
    else
        adamw_kernel<bf16><<<grid, 256, 0, s>>>(table, norm_sq, max_norm, lr, beta1, beta2, eps, step);
    VLK_CHECK_LAUNCH("vlk_adamw_step");
    return VLK_OK;
}
