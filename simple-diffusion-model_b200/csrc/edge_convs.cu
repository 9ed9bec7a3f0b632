// The two network-edge convolutions (models/U_Net.py:55-66 first conv, :113-130 last conv) on CUDA cores.
// With 3 (or 6) input channels the implicit-GEMM kernel must pad K from 27 to 9 x 64 = 576, with 3 output channels it
// computes a 64-column tile for 3 columns: ~20x wasted tensor work and operand traffic (1.1 ms of a 44 ms evaluation at
// batch 256).  Both layers hold ~0.1 % of the network's FLOPs, so direct fp32 FMA kernels at HBM speed are the better fit:
//   first: fp32 NCHW image -> NHWC `dtype`  (fuses the layout/pad pass), bias + optional Swish
//   last : NHWC `dtype` -> fp32 NCHW, bias + optional tanh                     (the network's output edge)
#include "host_util.h"
#include "ptx.cuh"
#include "stream.cuh"
#include "sdm_b200.h"

using namespace b2;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ first conv
// Thread = two horizontally adjacent output pixels x 16 output channels; a warp = 64 consecutive pixels of one channel
// group, so the weight reads (shared memory, [k][Cout], float4) are warp-uniform broadcasts and each one feeds 8 FMAs.
// x: [N][Cin][H][W] fp32 (W even); wk: [Cin*9][Cout] fp32.
template <typename T, int CIN>
__global__ void conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias,
                                     T* __restrict__ y, long long ldy, int N, int H, int W, int Cout, int act) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float ws[];                       // [CIN*9][Cout]
    constexpr int K = CIN * 9;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) ws[i] = wk[i];
    __syncthreads();
    const int groups = Cout / 16;
    const int cg = threadIdx.x / 32;                    // channel group of this warp
    const int lane = threadIdx.x % 32;
    const long long pairs = (long long)N * H * (W / 2);
    for (long long q0 = (long long)blockIdx.x * 32; q0 < pairs; q0 += (long long)gridDim.x * 32) {
        const long long q = q0 + lane;
        if (q >= pairs || cg >= groups) continue;
        const int w = (int)(q % (W / 2)) * 2, h = (int)((q / (W / 2)) % H);
        const long long n = q / ((long long)(W / 2) * H);
        float acc[2][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc[0][j] = bias ? __ldg(bias + cg * 16 + j) : 0.f; acc[1][j] = acc[0][j]; }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = h + kh - 1;
                const bool rok = hh >= 0 && hh < H;
                const float* row = x + ((n * CIN + ci) * H + (rok ? hh : 0)) * W;
                float v[4];                              // input columns w-1 .. w+2 serve both pixels
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ww = w + i - 1;
                    v[i] = (rok && ww >= 0 && ww < W) ? __ldg(row + ww) : 0.f;
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4* wr = reinterpret_cast<const float4*>(ws + (ci * 9 + kh * 3 + kw) * Cout + cg * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 w4 = wr[j];
                        acc[0][4 * j] = fmaf(v[kw], w4.x, acc[0][4 * j]);         acc[0][4 * j + 1] = fmaf(v[kw], w4.y, acc[0][4 * j + 1]);
                        acc[0][4 * j + 2] = fmaf(v[kw], w4.z, acc[0][4 * j + 2]); acc[0][4 * j + 3] = fmaf(v[kw], w4.w, acc[0][4 * j + 3]);
                        acc[1][4 * j] = fmaf(v[kw + 1], w4.x, acc[1][4 * j]);         acc[1][4 * j + 1] = fmaf(v[kw + 1], w4.y, acc[1][4 * j + 1]);
                        acc[1][4 * j + 2] = fmaf(v[kw + 1], w4.z, acc[1][4 * j + 2]); acc[1][4 * j + 3] = fmaf(v[kw + 1], w4.w, acc[1][4 * j + 3]);
                    }
                }
            }
        constexpr int V = V16<T>::N;
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            if (act == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[px][j] = swish_t<sizeof(T) == 2>(acc[px][j]);
            }
            T* o = y + ((n * H + h) * W + w + px) * ldy + cg * 16;
#pragma unroll
            for (int j = 0; j < 16 / V; ++j) {
                float vv[V];
#pragma unroll
                for (int i = 0; i < V; ++i) vv[i] = acc[px][j * V + i];
                stg16(o + j * V, pack16<T>(vv));
            }
        }
    }
}

extern "C" int b2_conv3x3_first(const float* x, const float* w_kc, const float* bias, void* y, long long ldy, int N, int Cin, int H,
                                int W, int Cout, int act, int dtype, void* stream) {
    if (Cout % 16 || Cout > 512) return set_error("b2_conv3x3_first: Cout must be a multiple of 16, <= 512");
    if (Cin != 3 && Cin != 6) return set_error("b2_conv3x3_first: Cin must be 3 or 6");
    if (W % 2) return set_error("b2_conv3x3_first: W must be even");
    if (ldy % (dtype == 0 ? 8 : 4) || ((uintptr_t)y & 15)) return set_error("b2_conv3x3_first: output must be 16-byte aligned");
    const int threads = 32 * (Cout / 16);
    const size_t smem = (size_t)Cin * 9 * Cout * sizeof(float);
    const long long pairs = (long long)N * H * (W / 2);
    long long blocks = (pairs + 31) / 32;
    const long long cap = 16LL * device_sm_count();
    if (blocks > cap) blocks = cap;
#define LAUNCH_FIRST(T, CIN)                                                                                               \
    do {                                                                                                                   \
        auto kern = conv3x3_first_kernel<T, CIN>;                                                                          \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        B2_LAUNCH((kern), (int)blocks, threads, smem, stream, x, w_kc, bias, (T*)y, ldy, N, H, W, Cout, act);              \
    } while (0)
    if (dtype == 0) { if (Cin == 3) LAUNCH_FIRST(bf16, 3); else LAUNCH_FIRST(bf16, 6); }
    else { if (Cin == 3) LAUNCH_FIRST(float, 3); else LAUNCH_FIRST(float, 6); }
#undef LAUNCH_FIRST
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_conv3x3_first: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------ last conv
// Thread = one output pixel, all (<= 4) output channels; a warp = 32 consecutive pixels, so neighbouring threads share
// most of their 3x3 input neighbourhoods through L1.  x: NHWC `dtype` (row stride ldx); wt: [9][Cin][4] fp32 (unused
// output slots zero); y: [N][Cout][H][W] fp32.
template <typename T>
__global__ void conv3x3_last_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ wt, const float* __restrict__ bias,
                                    float* __restrict__ y, int N, int H, int W, int Cin, int Cout, int act) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float ws[];                       // [9][Cin][4]
    for (int i = threadIdx.x; i < 9 * Cin * 4; i += blockDim.x) ws[i] = wt[i];
    __syncthreads();
    constexpr int V = V16<T>::N;
    constexpr int PX = 4;                               // horizontally adjacent pixels per thread: one weight read feeds 16 FMAs
    const int wq = W / PX;
    const long long total = (long long)N * H * wq;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int w0 = (int)(p % wq) * PX, h = (int)((p / wq) % H);
        const long long n = p / ((long long)wq * H);
        float acc[PX][4];
#pragma unroll
        for (int i = 0; i < PX; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            const int hh = h + t / 3 - 1, dw = t % 3 - 1;
            if (hh < 0 || hh >= H) continue;
            const T* srow = x + ((n * H + hh) * W) * ldx;
            const float4* wr = reinterpret_cast<const float4*>(ws + (long long)t * Cin * 4);
#pragma unroll 1
            for (int c = 0; c < Cin; c += V) {
                float v[PX][V];
#pragma unroll
                for (int i = 0; i < PX; ++i) {
                    const int ww = w0 + i + dw;
                    if (ww >= 0 && ww < W) unpack16<T>(ldg16(srow + (long long)ww * ldx + c), v[i]);
                    else {
#pragma unroll
                        for (int k = 0; k < V; ++k) v[i][k] = 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const float4 w4 = wr[c + k];
#pragma unroll
                    for (int i = 0; i < PX; ++i) {
                        acc[i][0] = fmaf(v[i][k], w4.x, acc[i][0]); acc[i][1] = fmaf(v[i][k], w4.y, acc[i][1]);
                        acc[i][2] = fmaf(v[i][k], w4.z, acc[i][2]); acc[i][3] = fmaf(v[i][k], w4.w, acc[i][3]);
                    }
                }
            }
        }
        for (int co = 0; co < Cout; ++co) {
            const float bv = bias ? __ldg(bias + co) : 0.f;
            float4 o;
            float* op = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                float v = acc[i][co] + bv;
                op[i] = act == 2 ? tanhf(v) : v;
            }
            *reinterpret_cast<float4*>(y + ((n * Cout + co) * H + h) * W + w0) = o;
        }
    }
}

extern "C" int b2_conv3x3_last(const void* x, long long ldx, const float* w_tc4, const float* bias, float* y, int N, int H, int W,
                               int Cin, int Cout, int act, int dtype, void* stream) {
    const int V = dtype == 0 ? 8 : 4;
    if (Cout < 1 || Cout > 4) return set_error("b2_conv3x3_last: Cout must be 1..4");
    if (Cin % V || ldx % V || ((uintptr_t)x & 15)) return set_error("b2_conv3x3_last: input channels / stride must be 16-byte aligned");
    const size_t smem = (size_t)9 * Cin * 4 * sizeof(float);
    if (smem > 200 * 1024) return set_error("b2_conv3x3_last: Cin too large");
    if (W % 4 || ((uintptr_t)y & 15)) return set_error("b2_conv3x3_last: W must be a multiple of 4 and y 16-byte aligned");
    const long long total = (long long)N * H * (W / 4);
    long long blocks = (total + 127) / 128;
    const long long cap = 16LL * device_sm_count();
    if (blocks > cap) blocks = cap;
    if (dtype == 0) {
        auto kern = conv3x3_last_kernel<bf16>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        B2_LAUNCH((kern), (int)blocks, 128, smem, stream, (const bf16*)x, ldx, w_tc4, bias, y, N, H, W, Cin, Cout, act);
    } else {
        auto kern = conv3x3_last_kernel<float>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        B2_LAUNCH((kern), (int)blocks, 128, smem, stream, (const float*)x, ldx, w_tc4, bias, y, N, H, W, Cin, Cout, act);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("b2_conv3x3_last: %s", cudaGetErrorString(e));
    return 0;
}
