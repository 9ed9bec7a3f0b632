// Building blocks of the memory-bound NHWC streaming kernels: 16-byte vector access, fp32 <-> storage conversion and a
// two-deep software pipeline over pixel rows (the loads of group i+1 are in flight while group i is processed, so the
// MUFU / FP32 work of Swish and the normalisation overlaps the HBM stream instead of alternating with it).
#pragma once
#include "ptx.cuh"

namespace b2 {

template <typename T> struct V16 { static constexpr int N = 16 / sizeof(T); };

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& u, float (&f)[V16<T>::N]) {
    if constexpr (sizeof(T) == 4) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
}
// fp32 storage feeds kind::tf32 MMAs: round to nearest instead of letting the tensor core truncate.
template <typename T>
__device__ __forceinline__ uint4 pack16(const float (&f)[V16<T>::N]) {
    uint4 u;
    if constexpr (sizeof(T) == 4) {
        u.x = __float_as_uint(round_tf32(f[0])); u.y = __float_as_uint(round_tf32(f[1]));
        u.z = __float_as_uint(round_tf32(f[2])); u.w = __float_as_uint(round_tf32(f[3]));
    } else {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
    return u;
}
__device__ __forceinline__ void stg16(void* p, const uint4& u) { *reinterpret_cast<uint4*>(p) = u; }

// V consecutive per-channel fp32 values (gamma, beta, AdaGN scale ...) with 16-byte loads; p must be 16-byte aligned.
template <int V>
__device__ __forceinline__ void ldg_f32(const float* p, float (&f)[V]) {
#pragma unroll
    for (int i = 0; i < V / 4; ++i) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p) + i);
        f[4 * i] = x.x; f[4 * i + 1] = x.y; f[4 * i + 2] = x.z; f[4 * i + 3] = x.w;
    }
}
// GroupNorm mean / rstd of the V channels starting at c0 from the (sum, sum of squares) pairs of the conv epilogue.
// When a group holds at least V channels (the usual case) all V share one pair: two loads instead of 2*V.
template <int V>
__device__ __forceinline__ void gn_mean_rstd(const float* __restrict__ stats_n /* [groups][2] of image n */, int c0, int cpg,
                                             float inv_cnt, float eps, float (&mean)[V], float (&rstd)[V]) {
    if (cpg % V == 0) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats_n) + c0 / cpg);
        const float m = st.x * inv_cnt;
        const float r = rsqrtf(fmaxf(st.y * inv_cnt - m * m, 0.f) + eps);
#pragma unroll
        for (int j = 0; j < V; ++j) { mean[j] = m; rstd[j] = r; }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float2 st = __ldg(reinterpret_cast<const float2*>(stats_n) + (c0 + j) / cpg);
            const float m = st.x * inv_cnt;
            mean[j] = m; rstd[j] = rsqrtf(fmaxf(st.y * inv_cnt - m * m, 0.f) + eps);
        }
    }
}

// load(buf, p) issues the global loads of row group p; proc(buf, p) consumes them.  Groups are `step` rows apart.
template <typename Buf, typename Load, typename Proc>
__device__ __forceinline__ void pipelined_rows(long long p_begin, long long p_end, long long step, Load load, Proc proc) {
    Buf b0, b1;
    long long p = p_begin;
    if (p < p_end) load(b0, p);
    while (p < p_end) {
        long long pn = p + step;
        if (pn < p_end) load(b1, pn);
        proc(b0, p);
        p = pn;
        if (p >= p_end) break;
        pn = p + step;
        if (pn < p_end) load(b0, pn);
        proc(b1, p);
        p = pn;
    }
}

}  // namespace b2
