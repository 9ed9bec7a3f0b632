// Memory-bound kernels of the backward pass (the reference gets these from autograd): GroupNorm x AdaGN backward
// (two streaming passes; the per-image finalize is folded into the second), Swish / tanh backward with fused bias gradients, query-axis
// softmax backward, column sums, and the kernel-layout -> parameter-layout gradient unpack.
#include "host_util.h"
#include "ptx.cuh"
#include "stream.cuh"
#include "sdm_b200.h"

using namespace b2;
typedef __nv_bfloat16 bf16;

#define LAUNCH_CHECK(name)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess) return set_error(name ": %s", cudaGetErrorString(e_));             \
        return 0;                                                                                 \
    } while (0)

template <typename T>
__device__ __forceinline__ void ld16(const T* p, float (&f)[V16<T>::N]) { unpack16<T>(*reinterpret_cast<const uint4*>(p), f); }
template <typename T>
__device__ __forceinline__ void st16(T* p, const float (&f)[V16<T>::N]) { stg16(p, pack16<T>(f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float swish_grad(float z) { const float s = sigmoidf_(z); return s * (1.0f + z * (1.0f - s)); }


// Per-channel partial sums held by `threads = cv * k` threads (thread owns V channels of pixel-row `prow`) are summed
// over the k pixel rows in shared memory; one fp32 atomic per channel per CTA then reaches global memory (instead of
// one per thread, which serialises thousands of same-address atomics in L2).
template <int V>
__device__ __forceinline__ void block_colsum_atomic(const float (&v)[V], float* red, int C, int c0, int prow, int k,
                                                    float* __restrict__ dst) {
#pragma unroll
    for (int j = 0; j < V; ++j) red[prow * C + c0 + j] = v[j];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < k; ++r) acc += red[r * C + c];
        atomicAdd(dst + c, acc);
    }
    __syncthreads();
}

// Launch shape shared by the streaming passes: each thread owns one 16-byte channel vector for the whole kernel
// (so per-channel partial sums live in registers) and strides over the pixels of one slab of one image.
struct SlabLaunch { int threads, rows_per_block, slabs; };
// ctas_per_sm: resident CTAs per SM of the kernel being launched (its __launch_bounds__).  The grid is sized to (just under) a
// whole number of waves of sms * ctas_per_sm CTAs: N * slabs = 592 CTAs on 444 slots ran 1.33 waves, i.e. a third of the
// machine idled through the second one.
static SlabLaunch slab_launch(int N, int HW, int cv, int ctas_per_sm = 0) {
    SlabLaunch s;
    int k = 256 / cv; if (k < 1) k = 1;
    s.threads = cv * k;
    s.rows_per_block = k;
    const int max_slabs = (HW + 4 * k - 1) / (4 * k);
    int slabs;
    if (ctas_per_sm <= 0) {
        slabs = (4 * device_sm_count() + N - 1) / N;
    } else {
        // the FEWEST slabs whose waves are >= 92 % full (every CTA pays a per-image prologue: statistics, group terms), else the
        // fullest; 592 CTAs on 444 slots ran 1.33 waves
        const int slots = device_sm_count() * ctas_per_sm;
        int hi = (4 * slots) / N + 1;
        if (hi > max_slabs) hi = max_slabs;
        if (hi < 1) hi = 1;
        slabs = 1;
        double best = -1.0;
        for (int cand = 1; cand <= hi; ++cand) {
            const long long total = (long long)N * cand;
            const long long waves = (total + slots - 1) / slots;
            const double eff = (double)total / (double)(waves * slots);
            if (eff > best + 1e-9) { best = eff; slabs = cand; }
            if (eff >= 0.92) { slabs = cand; break; }
        }
    }
    if (slabs > max_slabs) slabs = max_slabs;
    s.slabs = slabs < 1 ? 1 : slabs;
    return s;
}

// ------------------------------------------------------------------------------------------------ AdaGN backward
// Forward (custom_layers.py:35-45, :240-245): y = swish(z); xh = (y - mean) * rstd; out = s*(gamma*xh + beta) + s.
// Pass 1: a1[n][c] = sum_p dout, a2[n][c] = sum_p dout * xh.  The loop only accumulates (sum d, sum d*y): the statistics enter
// once per CTA, a2 = rstd * (sum d*y - mean * sum d), so neither mean nor rstd occupies registers while streaming.
// U = independent 16-byte row loads per tensor per thread in flight (x2 by the software pipeline): the kernels are bound by
// memory-level parallelism, so registers are spent on loads in flight, not on per-channel constants.
template <typename T, int U, int OCC>
__global__ void __launch_bounds__(256, OCC)
adagn_bwd_reduce_kernel(const T* __restrict__ dout, long long ldd, const T* __restrict__ z, long long ldz,
                        const float* __restrict__ stats, float* __restrict__ a1, float* __restrict__ a2,
                        int HW, int C, int groups, float eps, int slabs, int rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int V = V16<T>::N;
    const int cv = C / V;
    const int n = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int c0 = (threadIdx.x % cv) * V, prow = threadIdx.x / cv;
    float s1[V], s2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    const int p_per = (HW + slabs - 1) / slabs;
    const int p0 = slab * p_per, p1 = min(HW, p0 + p_per);
    const long long base = (long long)n * HW;
    constexpr bool kFast = sizeof(T) == 2;
    struct Buf { uint4 d[U], z[U]; };
    const long long k = rows_per_block;
    auto load = [&](Buf& b, long long p) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = p + u * k;
            if (q < p1) { b.d[u] = ldg16(dout + (base + q) * ldd + c0); b.z[u] = ldg16(z + (base + q) * ldz + c0); }
        }
    };
    auto proc = [&](Buf& b, long long p) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * k < p1) {
                float d[V], zz[V];
                unpack16<T>(b.d[u], d);
                unpack16<T>(b.z[u], zz);
#pragma unroll
                for (int j = 0; j < V; ++j) { s1[j] += d[j]; s2[j] = fmaf(d[j], swish_t<kFast>(zz[j]), s2[j]); }
            }
        }
    };
    pipelined_rows<Buf>(p0 + prow, p1, U * k, load, proc);
    {
        const int cpg = C / groups;
        const float inv_cnt = 1.0f / ((float)cpg * (float)HW);
        float mean[V], rstd[V];
        gn_mean_rstd<V>(stats + (long long)n * groups * 2, c0, cpg, inv_cnt, eps, mean, rstd);
#pragma unroll
        for (int j = 0; j < V; ++j) s2[j] = rstd[j] * (s2[j] - mean[j] * s1[j]);
    }
    extern __shared__ float red[];
    block_colsum_atomic<V>(s1, red, C, c0, prow, rows_per_block, a1 + (long long)n * C);
    block_colsum_atomic<V>(s2, red, C, c0, prow, rows_per_block, a2 + (long long)n * C);
}

// Pass 2: dz = rstd * (s*gamma*dout - m1 - xh*m2) * swish'(z);  dbias[c] += sum dz.
// Every CTA first folds the per-channel sums of pass 1 into the group terms of its image,
//   m[n][g] = (sum_{c in g} s*gamma*a1, sum_{c in g} s*gamma*a2) / (cpg*HW)          (C loads from L2, shared-memory adds),
// and the first slab of each image also emits ds[n][c] += gamma*a2 + (beta+1)*a1, dgamma[c] += s*a2, dbeta[c] += s*a1 --
// a separate finalize launch per layer (100 latency-bound launches per backward pass) is not needed.
// With xh = (y - mean) * rstd the normalisation gradient is affine in (dout, y) per channel,
//   dy = A*dout + B + Cc*y,  A = rstd*s*gamma,  Cc = -rstd^2*m2,  B = -rstd*m1 - Cc*mean,
// so three constants per channel stay in registers instead of five.
template <typename T, int U, int OCC>
__global__ void __launch_bounds__(256, OCC)
adagn_bwd_apply_kernel(const T* __restrict__ dout, long long ldd, const T* __restrict__ z, long long ldz,
                       const float* __restrict__ stats, const float* __restrict__ a1,
                       const float* __restrict__ a2, const float* __restrict__ s, long long s_bstride,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ ds,
                       long long ds_bstride, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       T* __restrict__ dz, long long lddz, float* __restrict__ dbias, int HW, int C,
                       int groups, float eps, int slabs, int rows_per_block, int raw_sums) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float red[];
    float* gs = red;                                  // [groups][2]; the same buffer serves the dbias reduction at the end
    constexpr int V = V16<T>::N;
    const int cv = C / V;
    const int n = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int c0 = (threadIdx.x % cv) * V, prow = threadIdx.x / cv;
    const int cpg = C / groups;
    const float inv_cnt = 1.0f / ((float)cpg * (float)HW);
    // group terms: warp w folds groups w, w + #warps, ...; lanes stride over the group's channels, shuffle reduction (the
    // previous shared-memory atomics were a 32-way same-address pile-up per group and dominated the small deep layers)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int g = warp; g < groups; g += nwarps) {
            float t1 = 0.f, t2 = 0.f;
            // raw_sums: a2 arrives as sum_p dout * y from the producing GEMM's epilogue; a2 = rstd * (sum dout*y - mean * sum dout)
            float g_mean = 0.f, g_rstd = 1.f;
            if (raw_sums) {
                const float2 st = __ldg(reinterpret_cast<const float2*>(stats + (long long)n * groups * 2) + g);
                g_mean = st.x * inv_cnt;
                g_rstd = rsqrtf(fmaxf(st.y * inv_cnt - g_mean * g_mean, 0.f) + eps);
            }
            for (int j = lane; j < cpg; j += 32) {
                const int c = g * cpg + j;
                const float x1 = a1[(long long)n * C + c];
                float x2 = a2[(long long)n * C + c];
                if (raw_sums) x2 = g_rstd * (x2 - g_mean * x1);
                const float sc = __ldg(s + (long long)n * s_bstride + c), ga = __ldg(gamma + c);
                if (slab == 0) {
                    atomicAdd(ds + (long long)n * ds_bstride + c, ga * x2 + (__ldg(beta + c) + 1.0f) * x1);
                    atomicAdd(dgamma + c, sc * x2);
                    atomicAdd(dbeta + c, sc * x1);
                }
                t1 = fmaf(sc * ga, x1, t1);
                t2 = fmaf(sc * ga, x2, t2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o); }
            if (lane == 0) { gs[g * 2] = t1; gs[g * 2 + 1] = t2; }
        }
    }
    __syncthreads();
    float A[V], B[V], Cc[V], db[V];
    {
        float mean[V], rstd[V], sc[V], ga[V];
        gn_mean_rstd<V>(stats + (long long)n * groups * 2, c0, cpg, inv_cnt, eps, mean, rstd);
        ldg_f32<V>(s + (long long)n * s_bstride + c0, sc);
        ldg_f32<V>(gamma + c0, ga);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int g = (c0 + j) / cpg;
            const float m1 = gs[g * 2] * inv_cnt, m2 = gs[g * 2 + 1] * inv_cnt;
            A[j] = rstd[j] * sc[j] * ga[j];
            Cc[j] = -rstd[j] * rstd[j] * m2;
            B[j] = -rstd[j] * m1 - Cc[j] * mean[j];
            db[j] = 0.f;
        }
    }
    __syncthreads();                                  // gs is dead from here on: `red` may be reused
    const int p_per = (HW + slabs - 1) / slabs;
    const int p0 = slab * p_per, p1 = min(HW, p0 + p_per);
    const long long base = (long long)n * HW;
    constexpr bool kFast = sizeof(T) == 2;
    struct Buf { uint4 d[U], z[U]; };
    const long long k = rows_per_block;
    auto load = [&](Buf& b, long long p) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = p + u * k;
            if (q < p1) { b.d[u] = ldg16(dout + (base + q) * ldd + c0); b.z[u] = ldg16(z + (base + q) * ldz + c0); }
        }
    };
    auto proc = [&](Buf& b, long long p) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = p + u * k;
            if (q < p1) {
                float d[V], zz[V], o[V];
                unpack16<T>(b.d[u], d);
                unpack16<T>(b.z[u], zz);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float zv = zz[j];
                    const float sig = sigmoid_t<kFast>(zv);
                    const float dy = fmaf(A[j], d[j], fmaf(Cc[j], zv * sig, B[j]));
                    o[j] = dy * (sig * (1.0f + zv * (1.0f - sig)));
                    db[j] += o[j];
                }
                stg16(dz + (base + q) * lddz + c0, pack16<T>(o));
            }
        }
    };
    pipelined_rows<Buf>(p0 + prow, p1, U * k, load, proc);
    if (dbias) block_colsum_atomic<V>(db, red, C, c0, prow, rows_per_block, dbias);
}

// L2 blocking of the two passes (SDM_B200_BWD_L2_CHUNK_MB=<MiB>, default 0 = off): both passes read dout and z, so running them
// over groups of images whose (dout, z) fit in a share of the 126 MB L2 could let pass 2 hit L2 instead of HBM.  MEASURED on B200
// (tools/bench_adagn_bwd.py, profiles/r02b_adagn_bwd_l2_chunk.log): it does not -- 32 / 64 / 96 MiB chunks are all SLOWER than the
// unblocked passes (7.3 ms -> 11.5 / 11.0 / 9.9 ms per backward pass at 128x128, batch 32): the re-read still misses and the
// extra, smaller launches cost more than they save.  Kept as a knob only.
static int adagn_bwd_chunk_images(int N, long long bytes_per_image) {
    static const long long budget = [] {
        const char* e = getenv("SDM_B200_BWD_L2_CHUNK_MB");
        return (e ? atoll(e) : 0LL) * (1LL << 20);
    }();
    if (budget <= 0 || bytes_per_image <= 0) return N;
    if ((long long)N * bytes_per_image <= budget) return N;
    long long nc = budget / bytes_per_image;
    if (nc < 1) nc = 1;
    // equal-sized chunks: ceil(N / ceil(N / nc))
    const long long chunks = (N + nc - 1) / nc;
    nc = (N + chunks - 1) / chunks;
    return (int)nc;
}

template <typename T, int U, int OCC>
static void adagn_bwd_launch2(int grid, int threads, size_t red_bytes, cudaStream_t st, const void* d_c, long long ldd,
                              const void* z_c, long long ldz, const float* stats_c, float* a1, float* a2, const float* s_c,
                              long long s_bstride, const float* gamma, const float* beta, float* ds_c, long long ds_bstride,
                              float* dgamma, float* dbeta, void* dz_c, long long lddz, float* dbias, int HW, int C, int groups,
                              float eps, int slabs, int rows_per_block, int sums_ready) {
    if (!sums_ready)
    B2_LAUNCH((adagn_bwd_reduce_kernel<T, U, OCC>), grid, threads, red_bytes, st, (const T*)d_c, ldd, (const T*)z_c, ldz, stats_c, a1, a2, HW, C, groups, eps, slabs, rows_per_block);
    B2_LAUNCH((adagn_bwd_apply_kernel<T, U, OCC>), grid, threads, red_bytes, st, (const T*)d_c, ldd, (const T*)z_c, ldz, stats_c, a1, a2, s_c, s_bstride, gamma, beta, ds_c, ds_bstride, dgamma, dbeta, (T*)dz_c, lddz, dbias, HW, C, groups, eps, slabs, rows_per_block, sums_ready);
}
#define ADAGN_BWD_ARGS grid, threads, red_bytes, st, d_c, ldd, z_c, ldz, stats_c, a1, a2, s_c, s_bstride, gamma, beta, ds_c, ds_bstride, dgamma, dbeta, dz_c, lddz, dbias, HW, C, groups, eps, slabs, rows_per_block, sums_ready
template <typename T>
static void adagn_bwd_launch(int U, int occ, int grid, int threads, size_t red_bytes, cudaStream_t st, const void* d_c, long long ldd,
                             const void* z_c, long long ldz, const float* stats_c, float* a1, float* a2, const float* s_c,
                             long long s_bstride, const float* gamma, const float* beta, float* ds_c, long long ds_bstride,
                             float* dgamma, float* dbeta, void* dz_c, long long lddz, float* dbias, int HW, int C, int groups,
                             float eps, int slabs, int rows_per_block, int sums_ready) {
    if (U == 4) adagn_bwd_launch2<T, 4, 2>(ADAGN_BWD_ARGS);
    else if (occ == 3) adagn_bwd_launch2<T, 2, 3>(ADAGN_BWD_ARGS);
    else adagn_bwd_launch2<T, 2, 2>(ADAGN_BWD_ARGS);
}

extern "C" int b2_adagn_bwd_fused(const void* dout, long long ldd, const void* z, long long ldz, const float* stats,
                                  const float* gamma, const float* beta, const float* s, long long s_bstride, float* work,
                                  float* ds, long long ds_bstride, float* dgamma, float* dbeta, void* dz, long long lddz,
                                  float* dbias, int N, int HW, int C, int groups, float eps, int sums_ready, int dtype, void* stream);
extern "C" int b2_adagn_bwd(const void* dout, long long ldd, const void* z, long long ldz, const float* stats,
                            const float* gamma, const float* beta, const float* s, long long s_bstride, float* work,
                            float* ds, long long ds_bstride, float* dgamma, float* dbeta, void* dz, long long lddz,
                            float* dbias, int N, int HW, int C, int groups, float eps, int dtype, void* stream) {
    return b2_adagn_bwd_fused(dout, ldd, z, ldz, stats, gamma, beta, s, s_bstride, work, ds, ds_bstride, dgamma, dbeta, dz, lddz, dbias,
                              N, HW, C, groups, eps, 0, dtype, stream);
}
extern "C" int b2_adagn_bwd_fused(const void* dout, long long ldd, const void* z, long long ldz, const float* stats,
                                  const float* gamma, const float* beta, const float* s, long long s_bstride, float* work,
                                  float* ds, long long ds_bstride, float* dgamma, float* dbeta, void* dz, long long lddz,
                                  float* dbias, int N, int HW, int C, int groups, float eps, int sums_ready, int dtype, void* stream) {
    const int V = dtype == 0 ? 8 : 4;
    if (C % V || ldd % V || ldz % V || lddz % V) return set_error("b2_adagn_bwd: channel counts / strides must be 16-byte aligned");
    if (C % groups) return set_error("b2_adagn_bwd: C %% groups != 0");
    const int cv = C / V;
    if (cv > 1024) return set_error("b2_adagn_bwd: C too large");
    cudaStream_t st = (cudaStream_t)stream;
    const long long eb = dtype == 0 ? 2 : 4;
    const int chunk = adagn_bwd_chunk_images(N, 2LL * HW * C * eb);
    static const int u_env = [] { const char* e = getenv("SDM_B200_BWD_U"); return e ? atoi(e) : 0; }();
    static const int occ_env = [] { const char* e = getenv("SDM_B200_BWD_OCC"); return e ? atoi(e) : 0; }();
    for (int n0 = 0; n0 < N; n0 += chunk) {
        const int nc = N - n0 < chunk ? N - n0 : chunk;
        // work: [2][N][C] fp32 (a1, a2), zeroed by the caller
        float* a1 = work + (long long)n0 * C;
        float* a2 = work + (long long)N * C + (long long)n0 * C;
        const char* d_c = (const char*)dout + (long long)n0 * HW * ldd * eb;
        const char* z_c = (const char*)z + (long long)n0 * HW * ldz * eb;
        char* dz_c = (char*)dz + (long long)n0 * HW * lddz * eb;
        const float* stats_c = stats + (long long)n0 * groups * 2;
        const float* s_c = s + (long long)n0 * s_bstride;
        float* ds_c = ds ? ds + (long long)n0 * ds_bstride : nullptr;
        // big images: four loads per tensor in flight (2 CTAs / SM at <= 128 registers); otherwise the two-load form (3 CTAs / SM)
        const int k0 = 256 / cv > 0 ? 256 / cv : 1;
        // MEASURED (profiles/r02c_adagn_bwd_variants.log, 100 layers at 128x128 batch 32): U=4 / 2 CTAs per SM 5.35 ms,
        // U=2 / 2 CTAs 5.61 ms, U=2 / 3 CTAs 6.36 ms
        (void)k0;
        const int U = u_env == 2 ? 2 : 4;
        const int occ = U == 4 ? 2 : (occ_env == 3 ? 3 : 2);
        const SlabLaunch sl = slab_launch(nc, HW, cv, occ);
        size_t red_bytes = (size_t)sl.rows_per_block * C * sizeof(float);
        if (red_bytes < (size_t)groups * 2 * sizeof(float)) red_bytes = (size_t)groups * 2 * sizeof(float);
        if (dtype == 0)
            adagn_bwd_launch<bf16>(U, occ, nc * sl.slabs, sl.threads, red_bytes, st, d_c, ldd, z_c, ldz, stats_c, a1, a2, s_c, s_bstride, gamma, beta, ds_c, ds_bstride, dgamma, dbeta, dz_c, lddz, dbias, HW, C, groups, eps, sl.slabs, sl.rows_per_block, sums_ready);
        else
            adagn_bwd_launch<float>(U, occ, nc * sl.slabs, sl.threads, red_bytes, st, d_c, ldd, z_c, ldz, stats_c, a1, a2, s_c, s_bstride, gamma, beta, ds_c, ds_bstride, dgamma, dbeta, dz_c, lddz, dbias, HW, C, groups, eps, sl.slabs, sl.rows_per_block, sums_ready);
    }
    LAUNCH_CHECK("b2_adagn_bwd");
}

// ------------------------------------------------------------------------------------------------ activation fwd/bwd
// mode 0: y = swish(z) (training forward of the un-normalised convs keeps z); mode 1: dz = dy * swish'(z), dbias += sum dz;
// mode 2: dz = dy (identity), dbias += sum dz (last conv / Linear bias gradients).
template <typename T>
__global__ void act_kernel(int mode, const T* __restrict__ a, long long lda, const T* __restrict__ z, long long ldz,
                           T* __restrict__ out, long long ldo, float* __restrict__ dbias, long long rows, int C,
                           int rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int V = V16<T>::N;
    const int cv = C / V;
    const int c0 = (threadIdx.x % cv) * V, prow = threadIdx.x / cv;
    float db[V];
#pragma unroll
    for (int j = 0; j < V; ++j) db[j] = 0.f;
    constexpr int U = 2;
    constexpr bool kFast = sizeof(T) == 2;
    struct Buf { uint4 a[U], z[U]; };
    const long long stride = (long long)gridDim.x * rows_per_block;
    auto load = [&](Buf& b, long long r0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * stride;
            if (r < rows) {
                if (mode != 0) b.a[u] = ldg16(a + r * lda + c0);
                if (mode != 2) b.z[u] = ldg16(z + r * ldz + c0);
            }
        }
    };
    auto proc = [&](Buf& b, long long r0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * stride;
            if (r < rows) {
                float x[V], zz[V], o[V];
                if (mode == 0) {
                    unpack16<T>(b.z[u], zz);
#pragma unroll
                    for (int j = 0; j < V; ++j) o[j] = swish_t<kFast>(zz[j]);
                    stg16(out + r * ldo + c0, pack16<T>(o));
                } else if (mode == 1) {
                    unpack16<T>(b.a[u], x);
                    unpack16<T>(b.z[u], zz);
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const float sig = sigmoid_t<kFast>(zz[j]);
                        o[j] = x[j] * (sig * (1.0f + zz[j] * (1.0f - sig)));
                        db[j] += o[j];
                    }
                    stg16(out + r * ldo + c0, pack16<T>(o));
                } else {
                    unpack16<T>(b.a[u], x);
#pragma unroll
                    for (int j = 0; j < V; ++j) db[j] += x[j];
                }
            }
        }
    };
    pipelined_rows<Buf>((long long)blockIdx.x * rows_per_block + prow, rows, U * stride, load, proc);
    if (dbias && mode != 0) {
        extern __shared__ float red[];
        block_colsum_atomic<V>(db, red, C, c0, prow, rows_per_block, dbias);
    }
}
extern "C" int b2_act(int mode, const void* a, long long lda, const void* z, long long ldz, void* out, long long ldo,
                      float* dbias, long long rows, int C, int dtype, void* stream) {
    const int V = dtype == 0 ? 8 : 4;
    if (C % V || (a && lda % V) || (z && ldz % V) || (out && ldo % V)) return set_error("b2_act: channel counts / strides must be 16-byte aligned");
    const int cv = C / V;
    if (cv > 1024) return set_error("b2_act: C too large");
    int k = 256 / cv; if (k < 1) k = 1;
    long long blocks = (rows + 4LL * k - 1) / (4LL * k);
    const long long cap = 8LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (dtype == 0) B2_LAUNCH((act_kernel<bf16>), (int)blocks, cv * k, (size_t)k * C * sizeof(float), (cudaStream_t)stream, mode, (const bf16*)a, lda, (const bf16*)z, ldz, (bf16*)out, ldo, dbias, rows, C, k);
    else B2_LAUNCH((act_kernel<float>), (int)blocks, cv * k, (size_t)k * C * sizeof(float), (cudaStream_t)stream, mode, (const float*)a, lda, (const float*)z, ldz, (float*)out, ldo, dbias, rows, C, k);
    LAUNCH_CHECK("b2_act");
}

// fp32 elementwise helpers for the tiny embedding MLPs: mode 0 y = swish(z); mode 1 dz = dy * swish'(z); mode 2 d *= 1 - y^2 (tanh).
__global__ void f32_act_kernel(int mode, const float* __restrict__ a, const float* __restrict__ z, float* __restrict__ out, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) out[i] = swishf(z[i]);
        else if (mode == 1) out[i] = a[i] * swish_grad(z[i]);
        else out[i] = a[i] * (1.0f - z[i] * z[i]);
    }
}
extern "C" int b2_f32_act(int mode, const float* a, const float* z, float* out, long long n, void* stream) {
    long long blocks = (n + 255) / 256;
    const long long cap = 8LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    B2_LAUNCH((f32_act_kernel), (int)blocks, 256, 0, (cudaStream_t)stream, mode, a, z, out, n);
    LAUNCH_CHECK("b2_f32_act");
}

// ------------------------------------------------------------------------------------------------ softmax backward
// dS[b][i][j] = scale * P[i][j] * (dP[i][j] - sum_i' P[i'][j] dP[i'][j])   (softmax over the query axis i)
template <typename T>
__global__ void softmax_query_axis_bwd_kernel(const T* __restrict__ P, const float* __restrict__ dP, T* __restrict__ dS,
                                              int B, int Pq, int Pk, long long ldp, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)B * Pk;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Pk);
        const long long b = idx / Pk;
        const T* pc = P + b * Pq * ldp + j;
        const float* dc = dP + b * Pq * Pk + j;
        float dot = 0.f;
        for (int i = 0; i < Pq; ++i) dot = fmaf(__ldg(dc + (long long)i * Pk), (float)pc[(long long)i * ldp], dot);
        T* o = dS + b * Pq * ldp + j;
        for (int i = 0; i < Pq; ++i) {
            const float v = scale * (float)pc[(long long)i * ldp] * (__ldg(dc + (long long)i * Pk) - dot);
            if constexpr (sizeof(T) == 4) o[(long long)i * ldp] = round_tf32(v); else o[(long long)i * ldp] = __float2bfloat16(v);
        }
    }
}
extern "C" int b2_softmax_query_axis_bwd(const void* P, const float* dP, void* dS, int B, int Pq, int Pk, long long ldp, float scale,
                                         int dtype, void* stream) {
    long long blocks = ((long long)B * Pk + 127) / 128;
    const long long cap = 8LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (dtype == 0) B2_LAUNCH((softmax_query_axis_bwd_kernel<bf16>), (int)blocks, 128, 0, (cudaStream_t)stream, (const bf16*)P, dP, (bf16*)dS, B, Pq, Pk, ldp, scale);
    else B2_LAUNCH((softmax_query_axis_bwd_kernel<float>), (int)blocks, 128, 0, (cudaStream_t)stream, (const float*)P, dP, (float*)dS, B, Pq, Pk, ldp, scale);
    LAUNCH_CHECK("b2_softmax_query_axis_bwd");
}

// ------------------------------------------------------------------------------------------------ gradient unpack
// kind 0: packed [Cout][9][Cin_pad] -> grad [Cout][Cin][3][3];  kind 2: packed [4][Cout][4][Cin] -> grad [Cin][Cout][4][4].
__global__ void unpack_weight_grad_kernel(int kind, const float* __restrict__ packed, float* __restrict__ grad, int Cout, int Cin,
                                          int Cin_pad, int accumulate) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = kind == 0 ? (long long)Cout * Cin * 9 : (long long)Cin * Cout * 16;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float v;
        if (kind == 0) {
            const int tap = (int)(i % 9); const int ci = (int)((i / 9) % Cin); const int co = (int)(i / (9LL * Cin));
            v = __ldg(packed + ((long long)co * 9 + tap) * Cin_pad + ci);
        } else {
            const int kw = (int)(i % 4), kh = (int)((i / 4) % 4); const int co = (int)((i / 16) % Cout); const int ci = (int)(i / (16LL * Cout));
            const int a = (kh == 1 || kh == 3) ? 0 : 1, ti = (kh == 1 || kh == 2) ? 0 : 1;
            const int b = (kw == 1 || kw == 3) ? 0 : 1, tj = (kw == 1 || kw == 2) ? 0 : 1;
            v = __ldg(packed + ((((long long)(a * 2 + b) * Cout + co) * 4) + ti * 2 + tj) * Cin + ci);
        }
        grad[i] = accumulate ? grad[i] + v : v;
    }
}
extern "C" int b2_unpack_weight_grad(int kind, const float* packed, float* grad, int Cout, int Cin, int Cin_pad, int accumulate,
                                     void* stream) {
    if (kind != 0 && kind != 2) return set_error("b2_unpack_weight_grad: kind must be 0 or 2");
    const long long total = kind == 0 ? (long long)Cout * Cin * 9 : (long long)Cin * Cout * 16;
    long long blocks = (total + 255) / 256;
    const long long cap = 16LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    B2_LAUNCH((unpack_weight_grad_kernel), (int)blocks, 256, 0, (cudaStream_t)stream, kind, packed, grad, Cout, Cin, Cin_pad, accumulate);
    LAUNCH_CHECK("b2_unpack_weight_grad");
}

// out = a + b on NHWC views (merging the two consumers of a skip tensor in the backward pass).
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, long long lda, const T* __restrict__ b, long long ldb, T* __restrict__ out,
                           long long ldo, long long rows, int cv) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int V = V16<T>::N;
    const long long total = rows * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cv; const int c = (int)(i % cv) * V;
        float x[V], y[V];
        ld16<T>(a + r * lda + c, x);
        ld16<T>(b + r * ldb + c, y);
#pragma unroll
        for (int j = 0; j < V; ++j) x[j] += y[j];
        st16<T>(out + r * ldo + c, x);
    }
}
extern "C" int b2_add(const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo, long long rows, int C,
                      int dtype, void* stream) {
    const int V = dtype == 0 ? 8 : 4;
    if (C % V || lda % V || ldb % V || ldo % V) return set_error("b2_add: channel counts / strides must be 16-byte aligned");
    long long blocks = (rows * (C / V) + 255) / 256;
    const long long cap = 8LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (dtype == 0) B2_LAUNCH((add_kernel<bf16>), (int)blocks, 256, 0, (cudaStream_t)stream, (const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)out, ldo, rows, C / V);
    else B2_LAUNCH((add_kernel<float>), (int)blocks, 256, 0, (cudaStream_t)stream, (const float*)a, lda, (const float*)b, ldb, (float*)out, ldo, rows, C / V);
    LAUNCH_CHECK("b2_add");
}
