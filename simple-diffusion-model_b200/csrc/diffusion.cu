// Fused elementwise kernels of the diffusion process itself (fp32 NCHW image tensors):
// q(x_t | x_0) noising, the DDPM / DDIM / cold-diffusion update steps, MSE loss + gradient, Philox normals.
// Each replaces 6-15 ATen launches of the reference with ONE float4-vectorised, grid-strided pass.
#include "host_util.h"
#include "ptx.cuh"
#include "sdm_b200.h"

using namespace b2;

#define LAUNCH_CHECK(name)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess) return set_error(name ": %s", cudaGetErrorString(e_));             \
        return 0;                                                                                 \
    } while (0)

static inline int ew_grid(long long n_vec, int threads) {
    long long blocks = (n_vec + threads - 1) / threads;
    const long long cap = (long long)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10
struct Philox {
    uint32_t k0, k1;
    __device__ Philox(unsigned long long seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ uint4 operator()(unsigned long long ctr, unsigned long long stream_id) const {
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)stream_id, c3 = (uint32_t)(stream_id >> 32);
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// four standard normals for global vector index `vec` (elements 4*vec .. 4*vec+3): independent of the launch shape,
// so a batch-sharded sampler draws exactly the numbers the unsharded one would.
__device__ __forceinline__ float4 philox_normal4(const Philox& ph, unsigned long long vec, unsigned long long offset) {
    const uint4 r = ph(vec, offset);
    const float r0 = sqrtf(-2.0f * __logf(u01(r.x))), r1 = sqrtf(-2.0f * __logf(u01(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.28318530718f * u01(r.y), &s0, &c0);
    __sincosf(6.28318530718f * u01(r.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__global__ void philox_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset,
                                     long long first_elem) {
    pdl_launch_dependents();
    pdl_wait();
    const Philox ph(seed);
    const long long nv = (n + 3) / 4;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nv; v += (long long)gridDim.x * blockDim.x) {
        const float4 z = philox_normal4(ph, (unsigned long long)(first_elem / 4 + v), offset);
        const long long i = v * 4;
        if (i + 3 < n) *reinterpret_cast<float4*>(out + i) = z;
        else { const float t[4] = {z.x, z.y, z.z, z.w}; for (int j = 0; i + j < n; ++j) out[i + j] = t[j]; }
    }
}
extern "C" int b2_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long offset,
                                long long first_elem, void* stream) {
    if (first_elem % 4) return set_error("b2_philox_normal: first_elem must be a multiple of 4");
    B2_LAUNCH((philox_normal_kernel), ew_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream, out, n, seed, offset, first_elem);
    LAUNCH_CHECK("b2_philox_normal");
}

// ------------------------------------------------------------------------------------------------ q(x_t | x_0)
// degraders.py:51-59 (linear: gather from the cumprod table) / :70-82,96-104 (cosine: closed form).
__device__ __forceinline__ float cosine_abar(float t, float T) {
    const float half_pi = 1.5707963267948966f;
    const float a = cosf(((t / T + 0.008f) / 1.008f) * half_pi);
    const float b = cosf(((0.0f / T + 0.008f) / 1.008f) * half_pi);
    return (a * a) / (b * b);
}
// Linear table: a timestep outside [0, max_step] poisons its image with NaN (the reference's torch.gather raises on it; here the trainer's
// NaN check / the caller's isfinite test trips instead of an out-of-bounds table read going unnoticed).
__device__ __forceinline__ float abar_of(long long t, const float* __restrict__ abar_table, int max_step) {
    if (!abar_table) return cosine_abar((float)t, (float)max_step);          // closed form: defined for every t, like the reference
    if (t < 0 || t > (long long)max_step) return __int_as_float(0x7fc00000);
    return abar_table[t];
}
// kPhilox: eps is not read but drawn in-kernel (Philox4x32-10 keyed on the GLOBAL element index, so data-parallel ranks and
// batch shards draw disjoint slices of one stream); `offset` selects the draw (the optimisation step), and when offset_dev
// is given the value is read from device memory instead -- the captured train step keeps its Adam step count there, so a
// replayed CUDA graph draws fresh noise every step without a separate RNG launch.  eps_out (optional) receives the draw.
template <bool kPhilox>
__global__ void qsample_kernel(const float* __restrict__ img, const float* __restrict__ eps, float* __restrict__ out,
                               float* __restrict__ eps_out, const long long* __restrict__ steps, int steps_count,
                               const float* __restrict__ abar_table, int max_step, long long per_image, int vec_per_image_blocks,
                               unsigned long long seed, unsigned long long offset, const float* __restrict__ offset_dev,
                               long long first_elem) {
    pdl_launch_dependents();
    pdl_wait();
    const int n = blockIdx.x / vec_per_image_blocks, blk = blockIdx.x % vec_per_image_blocks;
    const long long t = steps[steps_count == 1 ? 0 : n];
    const float abar = abar_of(t, abar_table, max_step);
    const float ca = sqrtf(abar), cb = sqrtf(1.0f - abar);
    const long long base = (long long)n * per_image;
    const long long nv = per_image / 4;
    const Philox ph(seed);
    if constexpr (kPhilox) { if (offset_dev) offset = (unsigned long long)__ldg(offset_dev); } else { (void)offset; }
    for (long long v = blk * (long long)blockDim.x + threadIdx.x; v < nv; v += (long long)vec_per_image_blocks * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(img + base) + v);
        float4 e;
        if constexpr (kPhilox) {
            e = philox_normal4(ph, (unsigned long long)((first_elem + base) / 4 + v), offset);
            if (eps_out) reinterpret_cast<float4*>(eps_out + base)[v] = e;
        } else {
            e = __ldg(reinterpret_cast<const float4*>(eps + base) + v);
        }
        float4 o;
        o.x = ca * a.x + cb * e.x; o.y = ca * a.y + cb * e.y; o.z = ca * a.z + cb * e.z; o.w = ca * a.w + cb * e.w;
        reinterpret_cast<float4*>(out + base)[v] = o;
    }
    if constexpr (!kPhilox) {
        for (long long i = nv * 4 + blk * (long long)blockDim.x + threadIdx.x; i < per_image; i += (long long)vec_per_image_blocks * blockDim.x)
            out[base + i] = ca * img[base + i] + cb * eps[base + i];
    }
}
static int qsample_blocks(int N, long long per_image) {
    int bpi = (int)((per_image / 4 + 255) / 256);
    const int cap = (device_sm_count() * 16 + N - 1) / N;
    if (bpi > cap) bpi = cap;
    return bpi < 1 ? 1 : bpi;
}
extern "C" int b2_qsample(const float* img, const float* eps, float* out, const long long* steps, int steps_count,
                          const float* abar_table, int max_step, int N, long long per_image, void* stream) {
    if (steps_count != 1 && steps_count != N) return set_error("b2_qsample: steps must have 1 or N entries");
    if (per_image % 4) return set_error("b2_qsample: C*H*W must be a multiple of 4");
    const int bpi = qsample_blocks(N, per_image);
    B2_LAUNCH((qsample_kernel<false>), N * bpi, 256, 0, (cudaStream_t)stream, img, eps, out, (float*)nullptr, steps, steps_count,
              abar_table, max_step, per_image, bpi, 0ull, 0ull, (const float*)nullptr, 0ll);
    LAUNCH_CHECK("b2_qsample");
}
extern "C" int b2_qsample_philox(const float* img, float* out, float* eps_out, const long long* steps, int steps_count,
                                 const float* abar_table, int max_step, int N, long long per_image, unsigned long long seed,
                                 unsigned long long offset, const float* offset_dev, long long first_elem, void* stream) {
    if (steps_count != 1 && steps_count != N) return set_error("b2_qsample_philox: steps must have 1 or N entries");
    if (per_image % 4 || first_elem % 4) return set_error("b2_qsample_philox: C*H*W and first_elem must be multiples of 4");
    const int bpi = qsample_blocks(N, per_image);
    B2_LAUNCH((qsample_kernel<true>), N * bpi, 256, 0, (cudaStream_t)stream, img, (const float*)nullptr, out, eps_out, steps,
              steps_count, abar_table, max_step, per_image, bpi, seed, offset, offset_dev, first_elem);
    LAUNCH_CHECK("b2_qsample_philox");
}

// ------------------------------------------------------------------------------------------------ sampler updates
// DDIM (diffusion_sampling_algorithms.py:107-136): x0 = c_scale*(x - c_s*e);  x' = c_an*x0 + c_dir*e + sigma*noise
__global__ void ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ e, const float* __restrict__ noise,
                                 float* __restrict__ x_out, float* __restrict__ x0_out, long long n, float c_scale, float c_s,
                                 float c_an, float c_dir, float sigma, int last) {
    pdl_launch_dependents();
    pdl_wait();
    const long long nv = n / 4;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nv; v += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + v), b = __ldg(reinterpret_cast<const float4*>(e) + v);
        float4 x0;
        x0.x = c_scale * (a.x - c_s * b.x); x0.y = c_scale * (a.y - c_s * b.y);
        x0.z = c_scale * (a.z - c_s * b.z); x0.w = c_scale * (a.w - c_s * b.w);
        if (x0_out) reinterpret_cast<float4*>(x0_out)[v] = x0;
        if (!last) {
            float4 o;
            o.x = c_an * x0.x + c_dir * b.x; o.y = c_an * x0.y + c_dir * b.y; o.z = c_an * x0.z + c_dir * b.z; o.w = c_an * x0.w + c_dir * b.w;
            if (noise) {
                const float4 z = __ldg(reinterpret_cast<const float4*>(noise) + v);
                o.x += sigma * z.x; o.y += sigma * z.y; o.z += sigma * z.z; o.w += sigma * z.w;
            }
            reinterpret_cast<float4*>(x_out)[v] = o;
        }
    }
}
extern "C" int b2_ddim_step(const float* x_t, const float* eps_hat, const float* noise, float* x_out, float* x0_out,
                            long long n, float c_scale, float c_s, float c_an, float c_dir, float sigma, int last, void* stream) {
    if (n % 4) return set_error("b2_ddim_step: element count must be a multiple of 4");
    B2_LAUNCH((ddim_step_kernel), ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream, x_t, eps_hat, noise, x_out, x0_out, n, c_scale, c_s, c_an, c_dir, sigma, last);
    LAUNCH_CHECK("b2_ddim_step");
}

// DDPM (diffusion_sampling_algorithms.py:42-55): x' = scale1*(x - scale2*e) + sigma*z; z given, or Philox in-kernel, or none.
__global__ void ddpm_step_kernel(const float* __restrict__ x, const float* __restrict__ e, const float* __restrict__ z_in,
                                 float* __restrict__ out, long long n, float scale1, float scale2, float sigma, int use_philox,
                                 unsigned long long seed, unsigned long long offset, long long first_elem) {
    pdl_launch_dependents();
    pdl_wait();
    const Philox ph(seed);
    const long long nv = n / 4;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nv; v += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + v), b = __ldg(reinterpret_cast<const float4*>(e) + v);
        float4 o;
        o.x = scale1 * (a.x - scale2 * b.x); o.y = scale1 * (a.y - scale2 * b.y);
        o.z = scale1 * (a.z - scale2 * b.z); o.w = scale1 * (a.w - scale2 * b.w);
        if (z_in || use_philox) {
            const float4 z = z_in ? __ldg(reinterpret_cast<const float4*>(z_in) + v)
                                  : philox_normal4(ph, (unsigned long long)(first_elem / 4 + v), offset);
            o.x += sigma * z.x; o.y += sigma * z.y; o.z += sigma * z.z; o.w += sigma * z.w;
        }
        reinterpret_cast<float4*>(out)[v] = o;
    }
}
extern "C" int b2_ddpm_step(const float* x_t, const float* eps_hat, const float* z, float* out, long long n, float scale1,
                            float scale2, float sigma, int use_philox, unsigned long long seed, unsigned long long offset,
                            long long first_elem, void* stream) {
    if (n % 4 || first_elem % 4) return set_error("b2_ddpm_step: element counts must be multiples of 4");
    B2_LAUNCH((ddpm_step_kernel), ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream, x_t, eps_hat, z, out, n, scale1, scale2, sigma, use_philox, seed, offset, first_elem);
    LAUNCH_CHECK("b2_ddpm_step");
}

// Cold diffusion (diffusion_sampling_algorithms.py:193-208): x' = x - D(x0, t) + D(x0, t'),  D(x0, t) = a_t*x0 + b_t*noise
__global__ void cold_step_kernel(const float* __restrict__ x, const float* __restrict__ x0, const float* __restrict__ noise,
                                 float* __restrict__ out, long long n, float a_t, float b_t, float a_n, float b_n) {
    pdl_launch_dependents();
    pdl_wait();
    const long long nv = n / 4;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nv; v += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + v), r = __ldg(reinterpret_cast<const float4*>(x0) + v),
                     z = __ldg(reinterpret_cast<const float4*>(noise) + v);
        float4 o;
        o.x = a.x - (a_t * r.x + b_t * z.x) + (a_n * r.x + b_n * z.x);
        o.y = a.y - (a_t * r.y + b_t * z.y) + (a_n * r.y + b_n * z.y);
        o.z = a.z - (a_t * r.z + b_t * z.z) + (a_n * r.z + b_n * z.z);
        o.w = a.w - (a_t * r.w + b_t * z.w) + (a_n * r.w + b_n * z.w);
        reinterpret_cast<float4*>(out)[v] = o;
    }
}
extern "C" int b2_cold_step(const float* x_t, const float* x0_hat, const float* noise, float* out, long long n, float a_t,
                            float b_t, float a_n, float b_n, void* stream) {
    if (n % 4) return set_error("b2_cold_step: element count must be a multiple of 4");
    B2_LAUNCH((cold_step_kernel), ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream, x_t, x0_hat, noise, out, n, a_t, b_t, a_n, b_n);
    LAUNCH_CHECK("b2_cold_step");
}

// ------------------------------------------------------------------------------------------------ MSE loss + gradient
// loss += sum((p - t)^2) * inv_n ; grad = 2 (p - t) * inv_n * grad_scale        (train_diffusion.py:350)
// kPhilox: the target is the eps of b2_qsample_philox, re-drawn here from the same (seed, offset, element index) instead of
// being stored by the q-sample kernel and read back (eps-prediction, train_diffusion.py:336-350).
template <bool kPhilox>
__global__ void mse_loss_grad_kernel(const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ grad,
                                     float* __restrict__ loss, long long n, float inv_n, float grad_scale,
                                     unsigned long long seed, unsigned long long offset, const float* __restrict__ offset_dev,
                                     long long first_elem) {
    pdl_launch_dependents();
    pdl_wait();
    float acc = 0.f;
    const long long nv = n / 4;
    const float gs = 2.0f * inv_n * grad_scale;
    const Philox ph(seed);
    if constexpr (kPhilox) { if (offset_dev) offset = (unsigned long long)__ldg(offset_dev); } else { (void)offset; }
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nv; v += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p) + v);
        float4 b;
        if constexpr (kPhilox) b = philox_normal4(ph, (unsigned long long)(first_elem / 4 + v), offset);
        else b = __ldg(reinterpret_cast<const float4*>(t) + v);
        const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        if (grad) reinterpret_cast<float4*>(grad)[v] = make_float4(gs * d.x, gs * d.y, gs * d.z, gs * d.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += wsum[i];
        atomicAdd(loss, s * inv_n);
    }
}
extern "C" int b2_mse_loss_grad(const float* pred, const float* target, float* grad, float* loss, long long n, float grad_scale,
                                void* stream) {
    if (n % 4) return set_error("b2_mse_loss_grad: element count must be a multiple of 4");
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("b2_mse_loss_grad: memset: %s", cudaGetErrorString(e));
    B2_LAUNCH((mse_loss_grad_kernel<false>), ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream, pred, target, grad, loss, n,
              1.0f / (float)n, grad_scale, 0ull, 0ull, (const float*)nullptr, 0ll);
    LAUNCH_CHECK("b2_mse_loss_grad");
}
extern "C" int b2_mse_loss_grad_philox(const float* pred, float* grad, float* loss, long long n, float grad_scale,
                                       unsigned long long seed, unsigned long long offset, const float* offset_dev,
                                       long long first_elem, void* stream) {
    if (n % 4 || first_elem % 4) return set_error("b2_mse_loss_grad_philox: element count and first_elem must be multiples of 4");
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("b2_mse_loss_grad_philox: memset: %s", cudaGetErrorString(e));
    B2_LAUNCH((mse_loss_grad_kernel<true>), ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream, pred, (const float*)nullptr, grad,
              loss, n, 1.0f / (float)n, grad_scale, seed, offset, offset_dev, first_elem);
    LAUNCH_CHECK("b2_mse_loss_grad_philox");
}

// ------------------------------------------------------------------------------------------------ area resample
// F.interpolate(mode="area") == adaptive average pooling (train_SR_diffusion.py:321-328, generate_sr_images_diffusion.py:170-173):
// out[o] = mean of in[floor(o*I/O) .. ceil((o+1)*I/O)) per axis.  Integer down-scaling averages k x k windows, integer
// up-scaling replicates.  fp32 NCHW planes; one thread per output element (consecutive threads walk the output row).
__global__ void area_resample_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int H, int W, int OH, int OW) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = planes * OH * OW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(i % OW), oh = (int)((i / OW) % OH);
        const long long pl = i / ((long long)OW * OH);
        const int h0 = (int)(((long long)oh * H) / OH), h1 = (int)((((long long)oh + 1) * H + OH - 1) / OH);
        const int w0 = (int)(((long long)ow * W) / OW), w1 = (int)((((long long)ow + 1) * W + OW - 1) / OW);
        const float* src = x + pl * H * W;
        float acc = 0.f;
        for (int h = h0; h < h1; ++h)
            for (int w = w0; w < w1; ++w) acc += __ldg(src + (long long)h * W + w);
        y[i] = acc / (float)((h1 - h0) * (w1 - w0));
    }
}
extern "C" int b2_area_resample(const float* x, float* y, long long planes, int H, int W, int OH, int OW, void* stream) {
    if (H < 1 || W < 1 || OH < 1 || OW < 1) return set_error("b2_area_resample: bad sizes");
    B2_LAUNCH((area_resample_kernel), ew_grid(planes * OH * OW, 256), 256, 0, (cudaStream_t)stream, x, y, planes, H, W, OH, OW);
    LAUNCH_CHECK("b2_area_resample");
}
