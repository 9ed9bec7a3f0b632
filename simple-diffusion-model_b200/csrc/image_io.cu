// Image input / output edges of the hot path on the device: memory-bound byte <-> fp32 conversions that the
// reference does on the host with numpy / torchvision, one coalesced pass each.
//   * b2_u8_to_image   uint8 HWC (cv2 BGR) -> fp32 CHW in [-1, 1], optional per-image horizontal flip
//                      (custom_dataset/img_dataset.py:26-35, train_diffusion.py:301-314, generate_sr_images_diffusion.py:117-126)
//   * b2_flip_images   per-image horizontal flip of an fp32 NCHW batch (torchvision RandomHorizontalFlip, train_diffusion.py:312-314)
//   * b2_image_to_u8   fp32 CHW in [-1, 1] -> uint8 HWC (the cascade hand-off image that generate_sr_images_diffusion.py:106-126 takes)
//   * b2_image_grid_u8 fp32 NCHW batch -> one uint8 HWC picture: channel swap (BGR -> RGB), make_grid(nrow, padding, normalize,
//                      value_range) and save_image's quantisation (utils/utils.py:39-65)
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

static inline int io_grid(long long items, int threads) {
    long long blocks = (items + threads - 1) / threads;
    const long long cap = (long long)device_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// (x - 127.5) / 127.5 exactly as the reference computes it: in double, rounded to fp32 once (img.astype(float) ... .float()).
__device__ __forceinline__ float u8_to_unit(unsigned char v) { return (float)(((double)v - 127.5) / 127.5); }

// One thread per output pixel (n, h, w): reads its C bytes (contiguous), writes C floats; consecutive threads walk the output
// row, so every plane's stores are coalesced and the byte loads of a warp cover one contiguous 32*C-byte run.
__global__ void u8_to_image_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, const unsigned char* __restrict__ flip,
                                   int N, int H, int W, int C) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
        const int ws = (flip && flip[n]) ? W - 1 - w : w;
        const unsigned char* s = src + (((long long)n * H + h) * W + ws) * C;
        float* d = dst + (long long)n * C * H * W + (long long)h * W + w;
        for (int c = 0; c < C; ++c) d[(long long)c * H * W] = u8_to_unit(s[c]);
    }
}
extern "C" int b2_u8_to_image(const void* src_u8_nhwc, float* dst_nchw, const void* flip_flags, int N, int H, int W, int C, void* stream) {
    if (N < 1 || H < 1 || W < 1 || C < 1 || C > 16) return set_error("b2_u8_to_image: bad shape");
    B2_LAUNCH((u8_to_image_kernel), io_grid((long long)N * H * W, 256), 256, 0, (cudaStream_t)stream,
              (const unsigned char*)src_u8_nhwc, dst_nchw, (const unsigned char*)flip_flags, N, H, W, C);
    LAUNCH_CHECK("b2_u8_to_image");
}

__global__ void flip_images_kernel(const float* __restrict__ x, float* __restrict__ out, const unsigned char* __restrict__ flip,
                                   long long total, int W, long long per_image) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        const long long n = i / per_image;
        out[i] = flip[n] ? __ldg(x + i - w + (W - 1 - w)) : __ldg(x + i);
    }
}
extern "C" int b2_flip_images(const float* x_nchw, float* out_nchw, const void* flip_flags, int N, int C, int H, int W, void* stream) {
    if (!flip_flags) return set_error("b2_flip_images: flip flags required");
    const long long per_image = (long long)C * H * W;
    B2_LAUNCH((flip_images_kernel), io_grid(N * per_image, 256), 256, 0, (cudaStream_t)stream, x_nchw, out_nchw,
              (const unsigned char*)flip_flags, N * per_image, W, per_image);
    LAUNCH_CHECK("b2_flip_images");
}

// save_image's quantisation of a [0, 1] value: mul(255).add_(0.5).clamp_(0, 255).to(uint8)   (truncation)
__device__ __forceinline__ unsigned char unit_to_u8(float v01) {
    float q = __fadd_rn(__fmul_rn(v01, 255.0f), 0.5f);          // two roundings like mul_().add_(): never contracted to an FMA
    q = fminf(fmaxf(q, 0.0f), 255.0f);
    return (unsigned char)q;
}
// make_grid's normalisation: clamp to [lo, hi], then (v - lo) / max(hi - lo, 1e-5)
__device__ __forceinline__ float norm_range(float v, float lo, float span) { return (v - lo) / span; }

__global__ void image_to_u8_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, int N, int C, int H, int W,
                                   float lo, float hi, float span) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / ((long long)H * W), hw = i % ((long long)H * W);
        const float* s = x + n * C * H * W + hw;
        unsigned char* d = out + i * C;
        for (int c = 0; c < C; ++c) {
            const float v = fminf(fmaxf(__ldg(s + (long long)c * H * W), lo), hi);
            d[c] = unit_to_u8(norm_range(v, lo, span));
        }
    }
}
extern "C" int b2_image_to_u8(const float* x_nchw, void* out_u8_nhwc, int N, int C, int H, int W, float lo, float hi, void* stream) {
    if (N < 1 || C < 1 || C > 16) return set_error("b2_image_to_u8: bad shape");
    const float span = hi - lo > 1e-5f ? hi - lo : 1e-5f;
    B2_LAUNCH((image_to_u8_kernel), io_grid((long long)N * H * W, 256), 256, 0, (cudaStream_t)stream, x_nchw,
              (unsigned char*)out_u8_nhwc, N, C, H, W, lo, hi, span);
    LAUNCH_CHECK("b2_image_to_u8");
}

// One thread per grid pixel: finds its cell, reads the C source values (channel order reversed when swap_rb), writes C bytes.
__global__ void image_grid_u8_kernel(const float* __restrict__ x, unsigned char* __restrict__ grid, int N, int C, int H, int W,
                                     int xmaps, int GH, int GW, int pad, int swap_rb, float lo, float hi, float span) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)GH * GW;
    const int ch = H + pad, cw = W + pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int gx = (int)(i % GW), gy = (int)(i / GW);
        unsigned char* d = grid + i * C;
        const int cy = gy - pad >= 0 ? (gy - pad) / ch : -1;
        const int cx = gx - pad >= 0 ? (gx - pad) / cw : -1;
        const int iy = gy - pad - cy * ch, ix = gx - pad - cx * cw;
        const int n = cy * xmaps + cx;
        const bool inside = cy >= 0 && cx >= 0 && cx < xmaps && iy >= 0 && iy < H && ix >= 0 && ix < W && n < N;
        for (int c = 0; c < C; ++c) {
            unsigned char q = 0;                                            // pad_value 0 -> black
            if (inside) {
                const int cs = swap_rb ? C - 1 - c : c;
                const float v = fminf(fmaxf(__ldg(x + (((long long)n * C + cs) * H + iy) * W + ix), lo), hi);
                q = unit_to_u8(norm_range(v, lo, span));
            }
            d[c] = q;
        }
    }
}
extern "C" int b2_image_grid_u8(const float* x_nchw, void* grid_u8_hwc, int N, int C, int H, int W, int nrow, int padding,
                                int swap_rb, float lo, float hi, void* stream) {
    if (N < 1 || C < 1 || C > 16 || nrow < 1 || padding < 0) return set_error("b2_image_grid_u8: bad arguments");
    // torchvision.utils.make_grid: a single image is returned as is (no border); otherwise xmaps = min(nrow, N) cells per row,
    // every cell (H + padding) x (W + padding), plus one leading border of `padding`
    int xmaps = nrow < N ? nrow : N, pad = padding;
    int GH, GW;
    if (N == 1) { pad = 0; xmaps = 1; GH = H; GW = W; }
    else { const int ymaps = (N + xmaps - 1) / xmaps; GH = ymaps * (H + pad) + pad; GW = xmaps * (W + pad) + pad; }
    const float span = hi - lo > 1e-5f ? hi - lo : 1e-5f;
    B2_LAUNCH((image_grid_u8_kernel), io_grid((long long)GH * GW, 256), 256, 0, (cudaStream_t)stream, x_nchw,
              (unsigned char*)grid_u8_hwc, N, C, H, W, xmaps, GH, GW, pad, swap_rb, lo, hi, span);
    LAUNCH_CHECK("b2_image_grid_u8");
}
