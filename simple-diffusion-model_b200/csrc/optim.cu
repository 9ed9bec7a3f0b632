// Fused Adam over flat fp32 buffers (parameters, gradients, exp_avg, exp_avg_sq): one memory-bound pass of
// 28 bytes per parameter instead of torch's multi-tensor foreach over ~760 tensors (train_diffusion.py:214-218,361).
#include "host_util.h"
#include "ptx.cuh"
#include "sdm_b200.h"
#include <cuda_bf16.h>

using namespace b2;

// kG16: gradients arrive as bf16 (data-parallel transport format, b200/parallel.py: the 2.43 GB fp32 exchange halves); all
// arithmetic and both moments stay fp32.
template <bool kG16>
__global__ void adam_flat_kernel(float* __restrict__ p, const void* __restrict__ g_, float* __restrict__ m, float* __restrict__ v,
                                 long long n, float beta1, float omb1, float beta2, float omb2, float eps, float step_size,
                                 float inv_bc2_sqrt, float grad_scale, const float* __restrict__ dev_state,
                                 __nv_bfloat16* __restrict__ shadow) {
    pdl_launch_dependents();
    pdl_wait();
    if (dev_state) {          // CUDA-graph mode: the step-dependent scalars come from device memory (b2_adam_flat_graph)
        grad_scale = dev_state[2]; step_size = dev_state[3]; inv_bc2_sqrt = dev_state[4];
    }
    const long long nv = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg;
        if constexpr (kG16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(g_) + i);
            const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
            const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
            gg = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
            gg = __ldg(reinterpret_cast<const float4*>(g_) + i);
        }
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = reinterpret_cast<float*>(&pp); const float* ga = reinterpret_cast<const float*>(&gg);
        float* ma = reinterpret_cast<float*>(&mm); float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = ga[j] * grad_scale;
            ma[j] = beta1 * ma[j] + omb1 * gr;
            va[j] = beta2 * va[j] + omb2 * (gr * gr);
            const float denom = sqrtf(va[j]) * inv_bc2_sqrt + eps;
            pa[j] -= step_size * (ma[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        if (shadow) {           // bf16 copy consumed by the tensor-core kernels: saves a separate cast pass over the weights
            uint2 sh;
            *reinterpret_cast<__nv_bfloat162*>(&sh.x) = __floats2bfloat162_rn(pp.x, pp.y);
            *reinterpret_cast<__nv_bfloat162*>(&sh.y) = __floats2bfloat162_rn(pp.z, pp.w);
            reinterpret_cast<uint2*>(shadow)[i] = sh;
        }
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = nv * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gr = (kG16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g_)[i]) : reinterpret_cast<const float*>(g_)[i]) * grad_scale;
        m[i] = beta1 * m[i] + omb1 * gr;
        v[i] = beta2 * v[i] + omb2 * (gr * gr);
        p[i] -= step_size * (m[i] / (sqrtf(v[i]) * inv_bc2_sqrt + eps));
        if (shadow) shadow[i] = __float2bfloat16(p[i]);
    }
}

// betas arrive as doubles so that 1 - beta is rounded once, like torch's Python-side scalar (1 - 0.999f != 0.001f).
// step_size = lr / (1 - beta1^t), inv_bc2_sqrt = 1 / sqrt(1 - beta2^t)  (torch.optim.Adam's formulation).
extern "C" int b2_adam_flat(float* p, const float* g, float* m, float* v, long long n, double beta1, double beta2, float eps,
                            float step_size, float inv_bc2_sqrt, float grad_scale, void* shadow_bf16, void* stream) {
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)shadow_bf16) & 15) return set_error("b2_adam_flat: buffers must be 16-byte aligned");
    // 128-thread blocks (6 K registers): small enough to share an SM with a resident tensor-core CTA, so a bucket's update can
    // run on a second stream underneath the rest of the backward pass (b200/parallel.py)
    long long blocks = (n / 4 + 127) / 128;
    const long long cap = 32LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    B2_LAUNCH((adam_flat_kernel<false>), (int)blocks, 128, 0, (cudaStream_t)stream, p, (const void*)g, m, v, n, (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps, step_size, inv_bc2_sqrt, grad_scale, (const float*)nullptr, (__nv_bfloat16*)shadow_bf16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_adam_flat: %s", cudaGetErrorString(e));
    return 0;
}

// state[0] = steps taken so far, state[1] = learning rate, state[2] = gradient scale; this kernel advances the step and
// derives state[3] = lr / (1 - beta1^t), state[4] = 1 / sqrt(1 - beta2^t) in double precision (torch computes them on the host).
__global__ void adam_advance_kernel(float* state, double beta1, double beta2) {
    pdl_launch_dependents();
    pdl_wait();
    const double t = (double)state[0] + 1.0;
    state[0] = (float)t;
    state[3] = (float)((double)state[1] / (1.0 - pow(beta1, t)));
    state[4] = (float)(1.0 / sqrt(1.0 - pow(beta2, t)));
}

extern "C" int b2_adam_advance(float* state, double beta1, double beta2, void* stream) {
    if (!state) return set_error("b2_adam_advance: state must be a device float[8]");
    B2_LAUNCH((adam_advance_kernel), 1, 1, 0, (cudaStream_t)stream, state, beta1, beta2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_adam_advance: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int b2_adam_flat_graph(float* p, const float* g, float* m, float* v, long long n, double beta1, double beta2,
                                  float eps, float* state, void* shadow_bf16, int advance, void* stream) {
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)shadow_bf16) & 15) return set_error("b2_adam_flat_graph: buffers must be 16-byte aligned");
    if (!state) return set_error("b2_adam_flat_graph: state must be a device float[8]");
    long long blocks = (n / 4 + 127) / 128;
    const long long cap = 32LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (advance) B2_LAUNCH((adam_advance_kernel), 1, 1, 0, (cudaStream_t)stream, state, beta1, beta2);
    B2_LAUNCH((adam_flat_kernel<false>), (int)blocks, 128, 0, (cudaStream_t)stream, p, (const void*)g, m, v, n, (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps, 0.f, 0.f, 0.f, (const float*)state, (__nv_bfloat16*)shadow_bf16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_adam_flat_graph: %s", cudaGetErrorString(e));
    return 0;
}

// Same update with bf16 gradients (the data-parallel transport buffer).  state == NULL: host scalars; else device scalars.
extern "C" int b2_adam_flat_g16(float* p, const void* g_bf16, float* m, float* v, long long n, double beta1, double beta2, float eps,
                                float step_size, float inv_bc2_sqrt, float grad_scale, float* state, void* shadow_bf16, void* stream) {
    if (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)shadow_bf16) & 15 || ((uintptr_t)g_bf16 & 7))
        return set_error("b2_adam_flat_g16: buffers must be 16-byte (gradients 8-byte) aligned");
    long long blocks = (n / 4 + 127) / 128;
    const long long cap = 32LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    B2_LAUNCH((adam_flat_kernel<true>), (int)blocks, 128, 0, (cudaStream_t)stream, p, g_bf16, m, v, n, (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps, step_size, inv_bc2_sqrt, grad_scale, (const float*)state, (__nv_bfloat16*)shadow_bf16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_adam_flat_g16: %s", cudaGetErrorString(e));
    return 0;
}

// fp32 <-> bf16 casts of a flat range (gradient buckets entering / leaving the bf16 all-reduce).
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    const long long nv = n / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(src) + i);
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(x.x, x.y);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(x.z, x.w);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
    for (long long i = nv * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16(src[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = __bfloat162float(src[i]);
}
extern "C" int b2_cast_f32_bf16(const float* src, void* dst_bf16, long long n, int to_f32, void* stream) {
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = 16LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (to_f32) B2_LAUNCH((cast_bf16_f32_kernel), (int)blocks, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)dst_bf16, const_cast<float*>(src), n);
    else {
        if (((uintptr_t)src & 15) || ((uintptr_t)dst_bf16 & 7)) return set_error("b2_cast_f32_bf16: buffers must be 16 / 8-byte aligned");
        B2_LAUNCH((cast_f32_bf16_kernel), (int)blocks, 256, 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst_bf16, n);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_cast_f32_bf16: %s", cudaGetErrorString(e));
    return 0;
}
