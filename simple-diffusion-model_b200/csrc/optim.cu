// Fused Adam over flat fp32 buffers (parameters, gradients, exp_avg, exp_avg_sq): one memory-bound pass of
// 28 bytes per parameter instead of torch's multi-tensor foreach over ~760 tensors (train_diffusion.py:214-218,361).
#include "host_util.h"
#include "ptx.cuh"
#include "sdm_b200.h"
#include <cuda_bf16.h>

using namespace b2;

__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
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
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
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
        const float gr = g[i] * grad_scale;
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
    B2_LAUNCH((adam_flat_kernel), (int)blocks, 128, 0, (cudaStream_t)stream, p, g, m, v, n, (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps, step_size, inv_bc2_sqrt, grad_scale, nullptr, (__nv_bfloat16*)shadow_bf16);
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
    B2_LAUNCH((adam_flat_kernel), (int)blocks, 128, 0, (cudaStream_t)stream, p, g, m, v, n, (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps, 0.f, 0.f, 0.f, state, (__nv_bfloat16*)shadow_bf16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_adam_flat_graph: %s", cudaGetErrorString(e));
    return 0;
}
