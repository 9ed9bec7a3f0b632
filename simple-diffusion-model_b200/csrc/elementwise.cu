// Memory-bound kernels of the forward path: layout edges, weight packing, GroupNorm x AdaGN apply (+residual),
// query-axis softmax, batched transpose, sinusoidal embedding, small fp32 linear.  All are vectorised
// (16-byte accesses where the layout allows) and sized in multiples of the SM count.
#include "host_util.h"
#include <cstring>
#include <cstdlib>
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

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return round_tf32(v); }   // fp32 activations feed kind::tf32 MMAs
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16(v); }

// 16-byte vector of T <-> floats
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };
template <typename T>
__device__ __forceinline__ void load16(const T* p, float (&f)[Vec16<T>::N]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    if constexpr (sizeof(T) == 4) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
}
template <typename T>
__device__ __forceinline__ void store16(T* p, const float (&f)[Vec16<T>::N]) {
    uint4 u;
    if constexpr (sizeof(T) == 4) {
        u.x = __float_as_uint(round_tf32(f[0])); u.y = __float_as_uint(round_tf32(f[1]));
        u.z = __float_as_uint(round_tf32(f[2])); u.w = __float_as_uint(round_tf32(f[3]));
    } else {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
    *reinterpret_cast<uint4*>(p) = u;
}

static inline int grid_for(long long work_items, int threads, int per_sm = 8) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)device_sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ------------------------------------------------------------------------------------------------ layout edges
// x fp32 NCHW [N][C][H][W]  ->  y T NHWC [N][H][W][Cpad] (channels >= C zero-filled).
template <typename T>
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, T* __restrict__ y, int N, int C, int HW, int Cpad) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)N * HW * (Cpad / 8);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % (Cpad / 8));
        const long long pix = i / (Cpad / 8);
        const int n = (int)(pix / HW), p = (int)(pix % HW);
        T* o = y + pix * Cpad + cv * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            o[j] = from_f<T>(c < C ? __ldg(x + ((long long)n * C + c) * HW + p) : 0.f);
        }
    }
}
extern "C" int b2_nchw_to_nhwc_pad(const float* x, void* y, int N, int C, int H, int W, int Cpad, int dtype, void* stream) {
    if (Cpad % 8) return set_error("b2_nchw_to_nhwc_pad: Cpad must be a multiple of 8");
    const long long total = (long long)N * H * W * (Cpad / 8);
    const int g = grid_for(total, 256);
    if (dtype == 0) B2_LAUNCH((nchw_to_nhwc_pad_kernel<bf16>), g, 256, 0, (cudaStream_t)stream, x, (bf16*)y, N, C, H * W, Cpad);
    else B2_LAUNCH((nchw_to_nhwc_pad_kernel<float>), g, 256, 0, (cudaStream_t)stream, x, (float*)y, N, C, H * W, Cpad);
    LAUNCH_CHECK("b2_nchw_to_nhwc_pad");
}

// x T NHWC (ld) [N][H][W][C] -> y fp32 NCHW.
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ y, int N, int C, int HW) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)N * C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int c = (int)((i / HW) % C);
        const int n = (int)(i / ((long long)HW * C));
        y[i] = to_f<T>(x[((long long)n * HW + p) * ldx + c]);
    }
}
extern "C" int b2_nhwc_to_nchw(const void* x, long long ldx, float* y, int N, int C, int H, int W, int dtype, void* stream) {
    const int g = grid_for((long long)N * C * H * W, 256);
    if (dtype == 0) B2_LAUNCH((nhwc_to_nchw_kernel<bf16>), g, 256, 0, (cudaStream_t)stream, (const bf16*)x, ldx, y, N, C, H * W);
    else B2_LAUNCH((nhwc_to_nchw_kernel<float>), g, 256, 0, (cudaStream_t)stream, (const float*)x, ldx, y, N, C, H * W);
    LAUNCH_CHECK("b2_nhwc_to_nchw");
}

// Parity planes for the stride-2 conv: planes[pr][pc][n][i][j][:] = x[n][2i+pr][2j+pc][:].
template <typename T>
__global__ void space_to_depth2_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ planes, int N, int H, int W, int C) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int V = Vec16<T>::N;
    const int cv = C / V, OH = H / 2, OW = W / 2;
    const long long total = (long long)N * H * W * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        long long pix = i / cv;
        const int w = (int)(pix % W); pix /= W;
        const int h = (int)(pix % H);
        const int n = (int)(pix / H);
        const uint4 v = *reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * ldx + c * V);
        const int pl = (h & 1) * 2 + (w & 1);
        T* o = planes + ((((long long)pl * N + n) * OH + (h >> 1)) * OW + (w >> 1)) * C + c * V;
        *reinterpret_cast<uint4*>(o) = v;
    }
}
extern "C" int b2_space_to_depth2(const void* x, long long ldx, void* planes, int N, int H, int W, int C, int dtype, void* stream) {
    const int V = dtype == 0 ? 8 : 4;
    if (C % V || ldx % V || (H & 1) || (W & 1)) return set_error("b2_space_to_depth2: C/ld must be vector aligned, H/W even");
    const int g = grid_for((long long)N * H * W * (C / V), 256);
    if (dtype == 0) B2_LAUNCH((space_to_depth2_kernel<bf16>), g, 256, 0, (cudaStream_t)stream, (const bf16*)x, ldx, (bf16*)planes, N, H, W, C);
    else B2_LAUNCH((space_to_depth2_kernel<float>), g, 256, 0, (cudaStream_t)stream, (const float*)x, ldx, (float*)planes, N, H, W, C);
    LAUNCH_CHECK("b2_space_to_depth2");
}

// ------------------------------------------------------------------------------------------------ weight packing
template <typename T> __device__ __forceinline__ T pack_val(float v);
template <> __device__ __forceinline__ bf16 pack_val<bf16>(float v) { return __float2bfloat16(v); }
template <> __device__ __forceinline__ float pack_val<float>(float v) { return round_tf32(v); }

// Kernel layouts ("K" is the contraction side of the GEMM the weight feeds; k_pad zero-pads it to the 128-byte block):
// kind 0: Conv2d [Cout][Cin][3][3]          -> [Cout][9][k_pad >= Cin]                  forward (stride 1 and 2)
// kind 1: Conv2d [Cout][Cin][3][3]          -> [Cin][9 flipped][k_pad >= Cout]          data gradient, stride 1
// kind 2: ConvTranspose2d [Cin][Cout][4][4] -> [4 parities][Cout][4 taps][Cin]          forward
// kind 3: Linear [rows=Cout][cols=Cin]      -> [rows][k_pad >= cols]                    forward
// kind 4: Linear [rows=Cout][cols=Cin]      -> [cols][k_pad >= rows] (transposed)       data gradient
// kind 5: Conv2d [Cout][Cin][3][3]          -> [4 parities][Cin][4 taps, zero padded][Cout]   data gradient, stride 2
// kind 6: ConvTranspose2d [Cin][Cout][4][4] -> [Cin][16][Cout]                          data gradient
template <typename T>
__device__ __forceinline__ T pack_weight_elem(int kind, const float* __restrict__ w, int Cout, int Cin, int k_pad, long long i) {
    float v = 0.f;
    if (kind == 0) {
        const int ci = (int)(i % k_pad); const int tap = (int)((i / k_pad) % 9); const int co = (int)(i / (9LL * k_pad));
        if (ci < Cin) v = __ldg(w + ((long long)co * Cin + ci) * 9 + tap);
    } else if (kind == 1) {
        const int co = (int)(i % k_pad); const int tap = (int)((i / k_pad) % 9); const int ci = (int)(i / (9LL * k_pad));
        if (co < Cout) v = __ldg(w + ((long long)co * Cin + ci) * 9 + (8 - tap));
    } else if (kind == 2) {
        const int ci = (int)(i % Cin); long long r = i / Cin;
        const int t = (int)(r % 4); r /= 4;
        const int co = (int)(r % Cout); const int g = (int)(r / Cout);
        const int a = g >> 1, b = g & 1, ti = t >> 1, tj = t & 1;
        const int kh = a == 0 ? (ti == 0 ? 1 : 3) : (ti == 0 ? 2 : 0);
        const int kw = b == 0 ? (tj == 0 ? 1 : 3) : (tj == 0 ? 2 : 0);
        v = __ldg(w + (((long long)ci * Cout + co) * 4 + kh) * 4 + kw);
    } else if (kind == 3) {
        const int c = (int)(i % k_pad); const int r = (int)(i / k_pad);
        if (c < Cin) v = __ldg(w + (long long)r * Cin + c);
    } else if (kind == 4) {
        const int r = (int)(i % k_pad); const int c = (int)(i / k_pad);
        if (r < Cout) v = __ldg(w + (long long)r * Cin + c);
    } else if (kind == 5) {
        const int co = (int)(i % Cout); long long r = i / Cout;
        const int t = (int)(r % 4); r /= 4;
        const int ci = (int)(r % Cin); const int g = (int)(r / Cin);
        const int a = g >> 1, b = g & 1, ti = t >> 1, tj = t & 1;
        const int kh = a == 0 ? (ti == 0 ? 1 : -1) : (ti == 0 ? 2 : 0);
        const int kw = b == 0 ? (tj == 0 ? 1 : -1) : (tj == 0 ? 2 : 0);
        if (kh >= 0 && kw >= 0) v = __ldg(w + ((long long)co * Cin + ci) * 9 + kh * 3 + kw);
    } else {
        const int co = (int)(i % Cout); const int t = (int)((i / Cout) % 16); const int ci = (int)(i / (16LL * Cout));
        v = __ldg(w + ((long long)ci * Cout + co) * 16 + t);
    }
    return pack_val<T>(v);
}
template <typename T>
__global__ void pack_weight_kernel(int kind, const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin, int k_pad,
                                   long long total) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = pack_weight_elem<T>(kind, w, Cout, Cin, k_pad, i);
}
// Several pack jobs in ONE launch (the ~80 kernel-layout weights a training step re-derives after the optimiser moved the
// masters are each a few microseconds of work: launching them one by one costs more latency than bandwidth).
// Job record (8 x int64): w, out, kind, Cout, Cin, k_pad, first global element, element count.
template <typename T>
__global__ void pack_weight_multi_kernel(const long long* __restrict__ jobs, int njobs, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int CHUNK = 4096;                          // consecutive elements per block: one job lookup per block
    __shared__ int job0;
    for (long long base = (long long)blockIdx.x * CHUNK; base < total; base += (long long)gridDim.x * CHUNK) {
        if (threadIdx.x == 0) {
            int lo = 0, hi = njobs - 1;                  // last job whose first element is <= base
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (jobs[mid * 8 + 6] <= base) lo = mid; else hi = mid - 1;
            }
            job0 = lo;
        }
        __syncthreads();
        int jb = job0;
        const long long end = base + CHUNK < total ? base + CHUNK : total;
        for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
            while (jb + 1 < njobs && jobs[(jb + 1) * 8 + 6] <= i) ++jb;
            const long long* j = jobs + jb * 8;
            reinterpret_cast<T*>(j[1])[i - j[6]] =
                pack_weight_elem<T>((int)j[2], reinterpret_cast<const float*>(j[0]), (int)j[3], (int)j[4], (int)j[5], i - j[6]);
        }
        __syncthreads();
    }
}
static long long pack_total(int kind, int Cout, int Cin, int k_pad) {
    switch (kind) {
        case 0: return (long long)Cout * 9 * k_pad;
        case 1: return (long long)Cin * 9 * k_pad;
        case 2: return 16LL * Cout * Cin;
        case 3: return (long long)Cout * k_pad;
        case 4: return (long long)Cin * k_pad;
        case 5: return 16LL * Cin * Cout;
        default: return 16LL * Cin * Cout;
    }
}
extern "C" int b2_pack_weight(int kind, const float* w, void* out, int Cout, int Cin, int k_pad, int dtype, void* stream) {
    if (kind < 0 || kind > 6) return set_error("b2_pack_weight: bad kind");
    const long long total = pack_total(kind, Cout, Cin, k_pad);
    const int g = grid_for(total, 256);
    if (dtype == 0) B2_LAUNCH((pack_weight_kernel<bf16>), g, 256, 0, (cudaStream_t)stream, kind, w, (bf16*)out, Cout, Cin, k_pad, total);
    else B2_LAUNCH((pack_weight_kernel<float>), g, 256, 0, (cudaStream_t)stream, kind, w, (float*)out, Cout, Cin, k_pad, total);
    LAUNCH_CHECK("b2_pack_weight");
}

extern "C" int b2_pack_weight_multi(const long long* jobs_dev, int njobs, long long total, int dtype, void* stream) {
    if (njobs < 1 || total < 1) return 0;
    const int g = grid_for((total + 15) / 16, 256);     // 4096-element chunks per 256-thread block
    if (dtype == 0) B2_LAUNCH((pack_weight_multi_kernel<bf16>), g, 256, 0, (cudaStream_t)stream, jobs_dev, njobs, total);
    else B2_LAUNCH((pack_weight_multi_kernel<float>), g, 256, 0, (cudaStream_t)stream, jobs_dev, njobs, total);
    LAUNCH_CHECK("b2_pack_weight_multi");
}

// Data-gradient weights of a 3x3/s1 conv straight from the bf16 "channels-last" copy the optimiser maintains:
// in [Cout][taps][Cin] -> out [Cin][taps flipped][Cout], taps = gridDim.z (9: a 3x3 weight; 1: a Linear weight).  64x64 tiles
// through shared memory, 128-byte rows on both sides.
__global__ void transpose_weight_cl_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int Cout, int Cin) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ bf16 tile[64][64 + 8];
    const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64, tap = blockIdx.z, taps = gridDim.z;
    const int t = threadIdx.x;                       // 256 threads
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * 256;                 // 512 16-byte vectors: 64 rows (co) x 8
        const int r = idx >> 3, v = idx & 7;
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(in + ((long long)(co0 + r) * taps + tap) * Cin + ci0 + v * 8));
        *reinterpret_cast<uint4*>(&tile[r][(v ^ ((r >> 3) & 7)) * 8]) = x;      // 16-byte chunks XOR-swizzled by row / 8
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * 256;                 // output rows (ci) x 8 vectors of 8 co
        const int r = idx >> 3, v = idx & 7;
        uint4 x;
        bf16* e = reinterpret_cast<bf16*>(&x);
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = tile[v * 8 + j][(((r >> 3) ^ v) << 3) + (r & 7)];    // conflict-free across v
        *reinterpret_cast<uint4*>(out + ((long long)(ci0 + r) * taps + (taps - 1 - tap)) * Cout + co0 + v * 8) = x;
    }
}
extern "C" int b2_transpose_weight_cl(const void* w_cl, void* out, int Cout, int Cin, void* stream) {
    if (Cout % 64 || Cin % 64) return set_error("b2_transpose_weight_cl: Cout and Cin must be multiples of 64");
    dim3 grid(Cin / 64, Cout / 64, 9);
    B2_LAUNCH((transpose_weight_cl_kernel), grid, 256, 0, (cudaStream_t)stream, (const bf16*)w_cl, (bf16*)out, Cout, Cin);
    LAUNCH_CHECK("b2_transpose_weight_cl");
}
extern "C" int b2_transpose_linear_weight(const void* w, void* out, int rows, int cols, void* stream) {
    if (rows % 64 || cols % 64) return set_error("b2_transpose_linear_weight: rows and cols must be multiples of 64");
    dim3 grid(cols / 64, rows / 64, 1);
    B2_LAUNCH((transpose_weight_cl_kernel), grid, 256, 0, (cudaStream_t)stream, (const bf16*)w, (bf16*)out, rows, cols);
    LAUNCH_CHECK("b2_transpose_linear_weight");
}

// Both kernel layouts of a ConvTranspose2d(4, stride 2, pad 1) weight from its bf16 copy w[Cin][Cout][4][4] (the 16 taps of
// one (ci, co) pair are one 32-byte sector): a CTA stages a 32 ci x 32 co tile in shared memory and writes
//   fwd   [4 parities][Cout][4 taps][Cin]   (kind 2; 64-byte runs along ci)     and / or
//   dgrad [Cin][16][Cout]                   (kind 6; 64-byte runs along co).
__global__ void pack_convT_bf16_kernel(const bf16* __restrict__ w, bf16* __restrict__ fwd, bf16* __restrict__ dgrad, int Cin, int Cout) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ uint32_t tile[32][32][9];             // [ci][co][8 tap pairs + 1 pad word]
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    const int t = threadIdx.x;                       // 256 threads
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = t + i * 256;                 // 2048 16-byte vectors: (ci, co, half)
        const int half = idx & 1, co = (idx >> 1) & 31, ci = idx >> 6;
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(w + ((long long)(ci0 + ci) * Cout + co0 + co) * 16 + half * 8));
        uint32_t* d = &tile[ci][co][half * 4];
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
    }
    __syncthreads();
    auto elem = [&](int ci, int co, int tap) -> uint32_t {          // the bf16 bits of w[ci][co][tap]
        const uint32_t pair = tile[ci][co][tap >> 1];
        return (tap & 1) ? (pair >> 16) : (pair & 0xffffu);
    };
    if (dgrad) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = t + i * 256;             // (ci, tap, 8-co vector)
            const int v = idx & 3, tap = (idx >> 2) & 15, ci = idx >> 6;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = elem(ci, v * 8 + 2 * j, tap) | (elem(ci, v * 8 + 2 * j + 1, tap) << 16);
            *reinterpret_cast<uint4*>(dgrad + ((long long)(ci0 + ci) * 16 + tap) * Cout + co0 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if (fwd) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = t + i * 256;             // (parity g, co, tap t4, 8-ci vector)
            const int v = idx & 3, t4 = (idx >> 2) & 3, co = (idx >> 4) & 31, g = idx >> 9;
            const int a = g >> 1, b = g & 1, ti = t4 >> 1, tj = t4 & 1;
            const int kh = a == 0 ? (ti == 0 ? 1 : 3) : (ti == 0 ? 2 : 0);
            const int kw = b == 0 ? (tj == 0 ? 1 : 3) : (tj == 0 ? 2 : 0);
            const int tap = kh * 4 + kw;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = elem(v * 8 + 2 * j, co, tap) | (elem(v * 8 + 2 * j + 1, co, tap) << 16);
            *reinterpret_cast<uint4*>(fwd + (((long long)g * Cout + co0 + co) * 4 + t4) * Cin + ci0 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}
extern "C" int b2_pack_convt_bf16(const void* w, void* fwd, void* dgrad, int Cin, int Cout, void* stream) {
    if (Cin % 32 || Cout % 32) return set_error("b2_pack_convt_bf16: Cin and Cout must be multiples of 32");
    dim3 grid(Cin / 32, Cout / 32, 1);
    B2_LAUNCH((pack_convT_bf16_kernel), grid, 256, 0, (cudaStream_t)stream, (const bf16*)w, (bf16*)fwd, (bf16*)dgrad, Cin, Cout);
    LAUNCH_CHECK("b2_pack_convt_bf16");
}

// ------------------------------------------------------------------------------------------------ GroupNorm x AdaGN apply
// out = s * (gamma * (y - mean) * rstd + beta) + s  (+ residual)        [custom_layers.py:35-45, :282-287]
// stats = per-(image, group) (sum, sum of squares) produced by the conv epilogue.  One CTA streams a slab of
// pixels of one image; the per-channel affine (a, b) is folded once into shared memory.
template <typename T, int U>
__global__ void adagn_apply_kernel(const T* __restrict__ y, long long ldy, const float* __restrict__ stats,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ s, long long s_bstride, const T* __restrict__ res, long long ldr,
                                   T* __restrict__ out, long long ldo, int HW, int C, int groups, float eps, int slabs,
                                   int rows_per_block, int pre_swish) {
    pdl_launch_dependents();
    pdl_wait();
    // each thread owns one 16-byte channel vector (its folded affine lives in registers) and walks the pixels of its
    // slab with U independent 16-byte loads in flight
    constexpr int V = V16<T>::N;
    const int cv = C / V;
    const int n = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int c0 = (threadIdx.x % cv) * V, prow = threadIdx.x / cv;
    const int cpg = C / groups;
    const float inv_cnt = 1.0f / ((float)cpg * (float)HW);
    float fa[V], fb[V];
    {
        float mean[V], rstd[V], sc[V], ga[V], be[V];
        gn_mean_rstd<V>(stats + (long long)n * groups * 2, c0, cpg, inv_cnt, eps, mean, rstd);
        ldg_f32<V>(s + (long long)n * s_bstride + c0, sc);
        ldg_f32<V>(gamma + c0, ga);
        ldg_f32<V>(beta + c0, be);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float g = ga[j] * rstd[j];
            fa[j] = sc[j] * g;
            fb[j] = sc[j] * (be[j] - g * mean[j]) + sc[j];
        }
    }
    const int p_per = (HW + slabs - 1) / slabs;
    const int p0 = slab * p_per, p1 = min(HW, p0 + p_per);
    const long long base = (long long)n * HW;
    constexpr bool kFast = sizeof(T) == 2;          // U row vectors in flight per pipeline half (U = 4 measured slower: registers cost occupancy)
    struct Buf { uint4 y[U], r[U]; };
    const long long k = rows_per_block;
    auto load = [&](Buf& b, long long p) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = p + u * k;
            if (q < p1) {
                b.y[u] = ldg16(y + (base + q) * ldy + c0);
                if (res) b.r[u] = ldg16(res + (base + q) * ldr + c0);
            }
        }
    };
    auto proc = [&](Buf& b, long long p) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = p + u * k;
            if (q < p1) {
                float v[V];
                unpack16<T>(b.y[u], v);
                if (pre_swish) {        // training: the conv stored its pre-activation z; Swish is applied here on the fly
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] = swish_t<kFast>(v[j]);
                }
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = fmaf(v[j], fa[j], fb[j]);
                if (res) {
                    float r[V];
                    unpack16<T>(b.r[u], r);
#pragma unroll
                    for (int j = 0; j < V; ++j) v[j] += r[j];
                }
                stg16(out + (base + q) * ldo + c0, pack16<T>(v));
            }
        }
    };
    pipelined_rows<Buf>(p0 + prow, p1, U * k, load, proc);
}
extern "C" int b2_adagn_apply(const void* y, long long ldy, const float* stats, const float* gamma, const float* beta,
                              const float* s, long long s_bstride, const void* residual, long long ldr, void* out,
                              long long ldo, int N, int HW, int C, int groups, float eps, int pre_swish, int dtype, void* stream) {
    const int V = dtype == 0 ? 8 : 4;
    if (C % V || ldy % V || ldo % V || (residual && ldr % V)) return set_error("b2_adagn_apply: channel counts / strides must be 16-byte aligned");
    if (C % groups) return set_error("b2_adagn_apply: C %% groups != 0");
    const int cv = C / V;
    if (cv > 1024) return set_error("b2_adagn_apply: C too large");
    const int sms = device_sm_count();
    int k = 256 / cv; if (k < 1) k = 1;
    const int max_slabs = (HW + 4 * k - 1) / (4 * k);
    // grid = the fewest slabs per image whose waves of (SMs x 3 resident CTAs: 80 registers, 256 threads) are >= 92 % full, up to
    // ~4 waves (same rule as the backward passes, csrc/backward.cu: a ragged last wave idles a third of the machine);
    // SDM_B200_APPLY_GRID=legacy restores ceil(8 * SMs / N)
    static const bool legacy_grid = [] { const char* e = getenv("SDM_B200_APPLY_GRID"); return e && !strcmp(e, "legacy"); }();
    int slabs = 1;
    if (legacy_grid) {
        slabs = (8 * sms + N - 1) / N;
    } else {
        const int slots = sms * 3;
        int hi = (4 * slots) / N + 1;
        if (hi > max_slabs) hi = max_slabs;
        if (hi < 1) hi = 1;
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
    if (slabs < 1) slabs = 1;
    if (dtype == 0)
        B2_LAUNCH((adagn_apply_kernel<bf16, 2>), N * slabs, cv * k, 0, (cudaStream_t)stream, (const bf16*)y, ldy, stats, gamma, beta, s, s_bstride,
                                                                                  (const bf16*)residual, ldr, (bf16*)out, ldo, HW, C, groups, eps, slabs, k, pre_swish);
    else
        B2_LAUNCH((adagn_apply_kernel<float, 2>), N * slabs, cv * k, 0, (cudaStream_t)stream, (const float*)y, ldy, stats, gamma, beta, s, s_bstride,
                                                                                   (const float*)residual, ldr, (float*)out, ldo, HW, C, groups, eps, slabs, k, pre_swish);
    LAUNCH_CHECK("b2_adagn_apply");
}

// All channels-last data-gradient transposes of a step in one launch.  Job record (6 x int64): in, out, Cout, Cin, first
// tile, tile count (tiles = Cin/64 x Cout/64 x 9 per weight).
__global__ void transpose_weight_cl_multi_kernel(const long long* __restrict__ jobs, int njobs) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ bf16 tile[64][64 + 8];
    int lo = 0, hi = njobs - 1;
    const long long b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid * 6 + 4] <= b) lo = mid; else hi = mid - 1;
    }
    const long long* j = jobs + lo * 6;
    const bf16* in = reinterpret_cast<const bf16*>(j[0]);
    bf16* out = reinterpret_cast<bf16*>(j[1]);
    const int Cout = (int)j[2], Cin = (int)j[3];
    int t = (int)(b - j[4]);
    const int ci0 = (t % (Cin / 64)) * 64;  t /= (Cin / 64);
    const int co0 = (t % (Cout / 64)) * 64; const int tap = t / (Cout / 64);
    const int tid = threadIdx.x;
    constexpr int taps = 9;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * 256;
        const int r = idx >> 3, v = idx & 7;
        *reinterpret_cast<uint4*>(&tile[r][(v ^ ((r >> 3) & 7)) * 8]) =
            __ldg(reinterpret_cast<const uint4*>(in + ((long long)(co0 + r) * taps + tap) * Cin + ci0 + v * 8));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * 256;
        const int r = idx >> 3, v = idx & 7;
        uint4 x;
        bf16* e = reinterpret_cast<bf16*>(&x);
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = tile[v * 8 + k][(((r >> 3) ^ v) << 3) + (r & 7)];
        *reinterpret_cast<uint4*>(out + ((long long)(ci0 + r) * taps + (taps - 1 - tap)) * Cout + co0 + v * 8) = x;
    }
}
extern "C" int b2_transpose_weight_cl_multi(const long long* jobs_dev, int njobs, long long total_tiles, void* stream) {
    if (njobs < 1 || total_tiles < 1) return 0;
    if (total_tiles > 0x7fffffffLL) return set_error("b2_transpose_weight_cl_multi: too many tiles");
    B2_LAUNCH((transpose_weight_cl_multi_kernel), (int)total_tiles, 256, 0, (cudaStream_t)stream, jobs_dev, njobs);
    LAUNCH_CHECK("b2_transpose_weight_cl_multi");
}

// GroupNorm statistics as a separate pass, for widths the conv epilogue does not fuse (fewer than 4 channels per
// group: C = 32 or 64 with the reference's 32 groups).  stats[n][g] += (sum, sum of squares) of y (or of Swish(y)).
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ y, long long ldy, float* __restrict__ stats, int HW, int C, int groups,
                                int slabs, int rows_per_block, int pre_swish) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float gsum[];                  // [groups][2]
    constexpr int V = V16<T>::N;
    const int cv = C / V;
    const int n = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    const int c0 = (threadIdx.x % cv) * V, prow = threadIdx.x / cv;
    const int cpg = C / groups;
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) gsum[i] = 0.f;
    __syncthreads();
    const int p_per = (HW + slabs - 1) / slabs;
    const int p0 = slab * p_per, p1 = min(HW, p0 + p_per);
    float s1[V], s2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    for (int p = p0 + prow; p < p1; p += rows_per_block) {
        float v[V];
        unpack16<T>(ldg16(y + ((long long)n * HW + p) * ldy + c0), v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float x = pre_swish ? swish_t<sizeof(T) == 2>(v[j]) : v[j];
            s1[j] += x; s2[j] = fmaf(x, x, s2[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
        atomicAdd(&gsum[((c0 + j) / cpg) * 2], s1[j]);
        atomicAdd(&gsum[((c0 + j) / cpg) * 2 + 1], s2[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) atomicAdd(stats + (long long)n * groups * 2 + i, gsum[i]);
}
// Deterministic (and batch-size invariant) variant for SDM_B200_DETERMINISTIC / b2_set_deterministic: one CTA per (image, block
// of CB channels) walks all pixels of its image in a fixed order; the per-thread partial sums are combined by fixed-order
// shared-memory passes and the (sum, sum of squares) pairs are STORED -- no atomics anywhere, so the statistics of an image do
// not depend on launch shape, batch size or timing (sharded sampling == unsharded sampling bit for bit).
template <typename T>
__global__ void gn_stats_det_kernel(const T* __restrict__ y, long long ldy, float* __restrict__ stats, int HW, int C, int groups,
                                    int CB, int rows_per_block, int pre_swish) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float red[];                   // [rows_per_block][CB] partials, then [2][CB] column sums
    constexpr int V = V16<T>::N;
    const int cvb = CB / V;
    const int blocks_per_image = C / CB;
    const int n = blockIdx.x / blocks_per_image, cb0 = (blockIdx.x % blocks_per_image) * CB;
    const int cl = (threadIdx.x % cvb) * V, prow = threadIdx.x / cvb;
    const int cpg = C / groups;
    float s1[V], s2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    for (int p = prow; p < HW; p += rows_per_block) {
        float v[V];
        unpack16<T>(ldg16(y + ((long long)n * HW + p) * ldy + cb0 + cl), v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float x = pre_swish ? swish_t<sizeof(T) == 2>(v[j]) : v[j];
            s1[j] += x; s2[j] = fmaf(x, x, s2[j]);
        }
    }
    float* csum = red + rows_per_block * CB;
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int j = 0; j < V; ++j) red[prow * CB + cl + j] = pass == 0 ? s1[j] : s2[j];
        __syncthreads();
        for (int c = threadIdx.x; c < CB; c += blockDim.x) {
            float a = 0.f;
            for (int r = 0; r < rows_per_block; ++r) a += red[r * CB + c];
            csum[pass * CB + c] = a;
        }
        __syncthreads();
    }
    const int g_local = CB / cpg;                     // groups owned by this CTA (CB is a multiple of cpg)
    for (int i = threadIdx.x; i < g_local * 2; i += blockDim.x) {
        const int g = i >> 1, which = i & 1;
        float a = 0.f;
        for (int c = 0; c < cpg; ++c) a += csum[which * CB + g * cpg + c];
        stats[((long long)n * groups + cb0 / cpg + g) * 2 + which] = a;
    }
}
namespace b2 {
int launch_gn_stats(const void* y, long long ldy, float* stats, int N, int HW, int C, int groups, int pre_swish, int dtype,
                    cudaStream_t st, bool deterministic) {
    const int V = dtype == 0 ? 8 : 4;
    if (C % V || ldy % V || C % groups) return set_error("gn_stats: channel count / stride must be 16-byte aligned and divisible by groups");
    if (deterministic) {
        const int cpg = C / groups;
        int CB = 64;                                  // channel block per CTA: a multiple of V and of the group width
        while (CB % cpg) CB *= 2;
        if (CB > C || C % CB) CB = C;
        const int cvb = CB / V;
        int k = 256 / cvb; if (k < 1) k = 1;
        const size_t smem = (size_t)(k + 2) * CB * sizeof(float);
        const int grid = N * (C / CB);
        if (dtype == 0) B2_LAUNCH((gn_stats_det_kernel<bf16>), grid, cvb * k, smem, st, (const bf16*)y, ldy, stats, HW, C, groups, CB, k, pre_swish);
        else B2_LAUNCH((gn_stats_det_kernel<float>), grid, cvb * k, smem, st, (const float*)y, ldy, stats, HW, C, groups, CB, k, pre_swish);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_error("gn_stats (deterministic): %s", cudaGetErrorString(e));
        return 0;
    }
    const int cv = C / V;
    int k = 256 / cv; if (k < 1) k = 1;
    int slabs = (4 * device_sm_count() + N - 1) / N;
    const int max_slabs = (HW + 4 * k - 1) / (4 * k);
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const size_t smem = (size_t)groups * 2 * sizeof(float);
    if (dtype == 0) B2_LAUNCH((gn_stats_kernel<bf16>), N * slabs, cv * k, smem, st, (const bf16*)y, ldy, stats, HW, C, groups, slabs, k, pre_swish);
    else B2_LAUNCH((gn_stats_kernel<float>), N * slabs, cv * k, smem, st, (const float*)y, ldy, stats, HW, C, groups, slabs, k, pre_swish);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("gn_stats: %s", cudaGetErrorString(e));
    return 0;
}
}  // namespace b2
extern "C" int b2_gn_stats(const void* y, long long ldy, float* stats, int N, int HW, int C, int groups, int pre_swish, int dtype,
                           void* stream) {
    return b2::launch_gn_stats(y, ldy, stats, N, HW, C, groups, pre_swish, dtype, (cudaStream_t)stream, b2::deterministic_mode());
}

// ------------------------------------------------------------------------------------------------ attention helpers
// Fix-up of the fused score kernel when a key row spans several 256-query tiles: tile t of row r holds
// exp2(s - m_t) and stats (m_t, z_t); the normalised probability is that value times exp2(m_t - m) / Z with
// m = max_t m_t, Z = sum_t z_t exp2(m_t - m).  One warp per row.
template <typename T>
__global__ void softmax_tiles_fixup_kernel(T* __restrict__ pt, long long ldp, const float* __restrict__ stats, long long rows,
                                           int P, int n_tiles, int tile_cols) {
    pdl_launch_dependents();
    pdl_wait();
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* st = stats + row * n_tiles * 2;
    float m = -INFINITY;
    for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, st[2 * t]);
    float z = 0.f;
    for (int t = 0; t < n_tiles; ++t) z += st[2 * t + 1] * exp2f(st[2 * t] - m);
    const float inv = 1.0f / z;
    T* o = pt + row * ldp;
    constexpr int V = 16 / (int)sizeof(T);      // one 16-byte vector per lane; a vector never straddles a tile (tile_cols % V == 0)
    if (P % V == 0 && ldp % V == 0 && tile_cols % V == 0 && (reinterpret_cast<uintptr_t>(pt) & 15) == 0) {
        for (int c = lane * V; c < P; c += 32 * V) {
            const float f = exp2f(st[2 * (c / tile_cols)] - m) * inv;
            uint4 raw = *reinterpret_cast<const uint4*>(o + c);
            T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
            for (int i = 0; i < V; ++i) e[i] = from_f<T>(to_f<T>(e[i]) * f);
            *reinterpret_cast<uint4*>(o + c) = raw;
        }
        return;
    }
    for (int c = lane; c < P; c += 32) {
        const float f = exp2f(st[2 * (c / tile_cols)] - m) * inv;
        o[c] = from_f<T>(to_f<T>(o[c]) * f);
    }
}
// dot[r][h] = sum_c a[r][h*hs + c] * b[r][h*hs + c], c < d  (softmax backward: sum_i P dP == sum_c V dV per key).
template <typename T>
__global__ void rowdot_kernel(const T* __restrict__ a, long long lda, const T* __restrict__ b, long long ldb, long long hs,
                              float* __restrict__ out, long long rows, int heads, int d) {
    pdl_launch_dependents();
    pdl_wait();
    const long long item = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= rows * heads) return;
    const long long r = item / heads;
    const int h = (int)(item % heads);
    const T* pa = a + r * lda + h * hs;
    const T* pb = b + r * ldb + h * hs;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) acc = fmaf(to_f<T>(pa[c]), to_f<T>(pb[c]), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[item] = acc;
}
namespace b2 {
int launch_softmax_fixup(void* pt, long long ldp, const float* stats, long long rows, int P, int n_tiles, int tile_cols, int dtype,
                         cudaStream_t st) {
    const long long threads = rows * 32;
    const int blocks = (int)((threads + 255) / 256);
    if (dtype == 0) B2_LAUNCH((softmax_tiles_fixup_kernel<bf16>), blocks, 256, 0, st, (bf16*)pt, ldp, stats, rows, P, n_tiles, tile_cols);
    else B2_LAUNCH((softmax_tiles_fixup_kernel<float>), blocks, 256, 0, st, (float*)pt, ldp, stats, rows, P, n_tiles, tile_cols);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("softmax fix-up: %s", cudaGetErrorString(e));
    return 0;
}
}  // namespace b2
extern "C" int b2_rowdot(const void* a, long long lda, const void* b, long long ldb, long long head_stride, float* out,
                         long long rows, int heads, int d, int dtype, void* stream) {
    const long long threads = rows * heads * 32;
    const int blocks = (int)((threads + 255) / 256);
    if (dtype == 0) B2_LAUNCH((rowdot_kernel<bf16>), blocks, 256, 0, (cudaStream_t)stream, (const bf16*)a, lda, (const bf16*)b, ldb, head_stride, out, rows, heads, d);
    else B2_LAUNCH((rowdot_kernel<float>), blocks, 256, 0, (cudaStream_t)stream, (const float*)a, lda, (const float*)b, ldb, head_stride, out, rows, heads, d);
    LAUNCH_CHECK("b2_rowdot");
}

// Query-axis softmax (custom_layers.py:147): S[b][i][j] fp32 (already scaled) -> P[b][i][j] = exp(S - max_i) / sum_i.
// One thread owns one key column j (coalesced across the warp), walking the query axis twice.
template <typename T>
__global__ void softmax_query_axis_kernel(const float* __restrict__ S, T* __restrict__ P, int B, int Pq, int Pk, long long ldp) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)B * Pk;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Pk);
        const long long b = idx / Pk;
        const float* col = S + b * Pq * Pk + j;
        float m = -INFINITY, z = 0.f;
        for (int i = 0; i < Pq; ++i) {
            const float v = col[(long long)i * Pk];
            const float mn = fmaxf(m, v);
            z = z * __expf(m - mn) + __expf(v - mn);
            m = mn;
        }
        const float inv = 1.0f / z;
        T* o = P + b * Pq * ldp + j;
        for (int i = 0; i < Pq; ++i) o[(long long)i * ldp] = from_f<T>(__expf(col[(long long)i * Pk] - m) * inv);
    }
}
extern "C" int b2_softmax_query_axis(const float* S, void* P, int B, int Pq, int Pk, long long ldp, int dtype, void* stream) {
    const int g = grid_for((long long)B * Pk, 128);
    if (dtype == 0) B2_LAUNCH((softmax_query_axis_kernel<bf16>), g, 128, 0, (cudaStream_t)stream, S, (bf16*)P, B, Pq, Pk, ldp);
    else B2_LAUNCH((softmax_query_axis_kernel<float>), g, 128, 0, (cudaStream_t)stream, S, (float*)P, B, Pq, Pk, ldp);
    LAUNCH_CHECK("b2_softmax_query_axis");
}

// out[b1][b2][c][r] = in[b1][b2][r][c]   (32x32 smem tiles); in strides in elements.
template <typename T>
__global__ void transpose_batched_kernel(const T* __restrict__ in, long long ld_in, long long in_s1, long long in_s2,
                                         T* __restrict__ out, long long ld_out, long long out_s1, long long out_s2,
                                         int R, int Ccols, int B1) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ T tile[32][33];
    const int b = blockIdx.z, b1 = b % B1, b2 = b / B1;
    const T* src = in + b1 * in_s1 + b2 * in_s2;
    T* dst = out + b1 * out_s1 + b2 * out_s2;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < Ccols) tile[i][threadIdx.x] = src[(long long)r * ld_in + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Ccols) dst[(long long)c * ld_out + r] = tile[threadIdx.x][i];
    }
}
extern "C" int b2_transpose_batched(const void* in, long long ld_in, long long in_s1, long long in_s2, void* out, long long ld_out,
                                    long long out_s1, long long out_s2, int R, int Ccols, int B1, int B2, int dtype, void* stream) {
    dim3 grid((Ccols + 31) / 32, (R + 31) / 32, B1 * B2), block(32, 8);
    if (grid.y > 65535 || grid.z > 65535) return set_error("b2_transpose_batched: grid too large");
    if (dtype == 0) B2_LAUNCH((transpose_batched_kernel<bf16>), grid, block, 0, (cudaStream_t)stream, (const bf16*)in, ld_in, in_s1, in_s2, (bf16*)out, ld_out, out_s1, out_s2, R, Ccols, B1);
    else B2_LAUNCH((transpose_batched_kernel<float>), grid, block, 0, (cudaStream_t)stream, (const float*)in, ld_in, in_s1, in_s2, (float*)out, ld_out, out_s1, out_s2, R, Ccols, B1);
    LAUNCH_CHECK("b2_transpose_batched");
}

// ------------------------------------------------------------------------------------------------ embedding + small linear
// out[b][:] = [sin(t_b f_k), cos(t_b f_k)], f_k = exp(-k ln(1e4)/(half-1))        (custom_layers.py:84-90)
__global__ void sinusoid_kernel(const long long* __restrict__ t, float* __restrict__ out, int B, int dim) {
    pdl_launch_dependents();
    pdl_wait();
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, k = i % half;
    const float f = expf((float)k * -(logf(10000.0f) / (float)(half - 1)));
    const float a = (float)t[b] * f;
    out[(long long)b * dim + k] = sinf(a);
    out[(long long)b * dim + half + k] = cosf(a);
}
extern "C" int b2_sinusoid_embedding(const long long* t, float* out, int B, int dim, void* stream) {
    if (dim < 4 || dim % 2) return set_error("b2_sinusoid_embedding: time_dim must be even and >= 4");
    const int n = B * (dim / 2);
    B2_LAUNCH((sinusoid_kernel), (n + 127) / 128, 128, 0, (cudaStream_t)stream, t, out, B, dim);
    LAUNCH_CHECK("b2_sinusoid_embedding");
}

// Small fp32 GEMM on CUDA cores for the embedding MLPs and the AdaGN scale vectors (M = batch, tiny):
//   C[M][N] (+)= op(A) . op(B) (+ bias[N]) (Swish).  ta: A is [K][M]; tb == 0: B is [N][K] (Linear weight), tb == 1: B is [K][N].
__global__ void small_gemm_kernel(const float* __restrict__ A, long long lda, int ta, const float* __restrict__ Bm, long long ldb, int tb,
                                  float* __restrict__ C, long long ldc, int M, int N, int K, const float* __restrict__ bias, int act,
                                  int accumulate, int k_per_split) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float As[32][33], Bs[32][33];
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // split-K (gridDim.z > 1): each z-slice reduces its K range and adds atomically into a zeroed / accumulated C
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);
    const bool split = gridDim.z > 1;
    for (int k0 = k_begin; k0 < k_end; k0 += 32) {
        for (int i = ty; i < 32; i += 8) {
            // As[i][tx] = A[m0+i][k0+tx]
            // tx always walks the contiguous dimension of the operand in memory (coalesced either way)
            if (ta) {
                const int k = k0 + i, m = m0 + tx;
                As[tx][i] = (m < M && k < k_end) ? A[(long long)k * lda + m] : 0.f;
            } else {
                const int m = m0 + i, k = k0 + tx;
                As[i][tx] = (m < M && k < k_end) ? A[(long long)m * lda + k] : 0.f;
            }
            if (tb) {
                const int k = k0 + i, n = n0 + tx;
                Bs[tx][i] = (n < N && k < k_end) ? Bm[(long long)k * ldb + n] : 0.f;
            } else {
                const int n = n0 + i, k = k0 + tx;
                Bs[i][tx] = (n < N && k < k_end) ? Bm[(long long)n * ldb + k] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const float b = Bs[tx][kk];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[ty + 8 * r][kk], b, acc[r]);
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    if (n < N) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = m0 + ty + 8 * r;
            if (m < M) {
                float v = acc[r] + ((bias && blockIdx.z == 0) ? bias[n] : 0.f);
                float* o = C + (long long)m * ldc + n;
                if (split) { atomicAdd(o, v); continue; }
                if (act == 1) v = swishf(v);
                *o = accumulate ? (*o + v) : v;
            }
        }
    }
}
extern "C" int b2_small_gemm(const float* A, long long lda, int ta, const float* B, long long ldb, int tb, float* C, long long ldc,
                             int M, int N, int K, const float* bias, int act, int accumulate, void* stream) {
    dim3 grid((N + 31) / 32, (M + 31) / 32), block(32, 8);
    // long contractions with a tiny output (d emb = ds_all . W_all: K = sum of all AdaGN widths) would run on a couple of
    // CTAs: split K across the grid instead
    int splits = 1;
    const long long tiles = (long long)grid.x * grid.y;
    if (act == 0 && K >= 2048 && tiles < 2LL * device_sm_count()) {
        splits = (int)((2LL * device_sm_count() + tiles - 1) / tiles);
        const int max_splits = (K + 255) / 256;
        if (splits > max_splits) splits = max_splits;
    }
    int k_per_split = (((K + splits - 1) / splits) + 31) / 32 * 32;
    splits = (K + k_per_split - 1) / k_per_split;
    if (splits > 1 && !accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, (cudaStream_t)stream);
        if (e != cudaSuccess) return set_error("b2_small_gemm: memset: %s", cudaGetErrorString(e));
    }
    grid.z = splits;
    B2_LAUNCH((small_gemm_kernel), grid, block, 0, (cudaStream_t)stream, A, lda, ta, B, ldb, tb, C, ldc, M, N, K, bias, act, accumulate, k_per_split);
    LAUNCH_CHECK("b2_small_gemm");
}
