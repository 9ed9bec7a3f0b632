// Standalone GPU check of the tcgen05 implicit-GEMM C-ABI against a CPU double-precision reference.
// Build: see Makefile target `tools`. Run on a B200: ./build/test_igemm [perf]
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "sdm_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define B2(x) do { int r_ = (x); if (r_) { printf("b2 error: %s (%s:%d)\n", b2_last_error(), __FILE__, __LINE__); exit(3); } } while (0)

static uint32_t rng_state = 12345;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f; }
static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

struct Buf {
    void* d = nullptr; size_t n = 0; int dtype = 0;
    std::vector<float> h;   // values as the GPU sees them (rounded)
    void alloc(size_t n_, int dtype_, bool random, float scale = 1.f) {
        n = n_; dtype = dtype_; h.resize(n);
        for (size_t i = 0; i < n; ++i) { float v = random ? frand() * scale : 0.f; h[i] = dtype == 0 ? bf16r(v) : tf32r(v); }
        upload();
    }
    void upload() {
        size_t eb = dtype == 0 ? 2 : 4;
        if (!d) CK(cudaMalloc(&d, n * eb + 256));
        if (dtype == 0) { std::vector<__nv_bfloat16> t(n); for (size_t i = 0; i < n; ++i) t[i] = __float2bfloat16(h[i]); CK(cudaMemcpy(d, t.data(), n * 2, cudaMemcpyHostToDevice)); }
        else CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    }
    void download() {
        if (dtype == 0) { std::vector<__nv_bfloat16> t(n); CK(cudaMemcpy(t.data(), d, n * 2, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < n; ++i) h[i] = __bfloat162float(t[i]); }
        else CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost));
    }
    void free_() { if (d) cudaFree(d); d = nullptr; }
};

static double swish(double x) { return x / (1.0 + exp(-x)); }
static int g_fail = 0;
static void report(const char* name, double maxerr, double tol) {
    printf("%-58s max_err %.3e  %s\n", name, maxerr, maxerr <= tol ? "OK" : "FAIL");
    if (!(maxerr <= tol)) g_fail++;
    fflush(stdout);
}

// conv reference on NHWC; w OIHW-like: w[co][kh][kw][ci]
static void test_conv(int mode, int dtype, int N, int H, int W, int Cin, int Cout, bool use_res, bool use_gn, int act) {
    // mode 0: s1; mode 1: s2 (H, W = input dims here, even); mode 2: convT (H, W input dims)
    const int OH = mode == 0 ? H : (mode == 1 ? H / 2 : 2 * H), OW = mode == 0 ? W : (mode == 1 ? W / 2 : 2 * W);
    const int KS = mode == 2 ? 4 : 3;
    Buf x, wp, y, res; std::vector<float> w((size_t)Cout * KS * KS * Cin), bias(Cout);
    x.alloc((size_t)N * H * W * Cin, dtype, true, 2.f);
    for (auto& v : w) { float t = frand() * 0.2f; v = dtype == 0 ? bf16r(t) : tf32r(t); }
    for (auto& v : bias) v = frand();
    // pack
    if (mode != 2) {
        wp.alloc(w.size(), dtype, false);
        wp.h = w;   // [co][kh*3+kw][ci] already
        wp.upload();
    } else {
        // logical w[co][kh][kw][ci]; packed [g=(a,b)][co][t=(i,j)][ci]; a=0: kh = {1,3}; a=1: kh = {2,0}
        wp.alloc((size_t)4 * Cout * 4 * Cin, dtype, false);
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int co = 0; co < Cout; ++co)
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                int kh = a == 0 ? (i == 0 ? 1 : 3) : (i == 0 ? 2 : 0);
                int kw = b == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 2 : 0);
                for (int ci = 0; ci < Cin; ++ci)
                    wp.h[((((size_t)(a * 2 + b) * Cout + co) * 4) + i * 2 + j) * Cin + ci] = w[(((size_t)co * 4 + kh) * 4 + kw) * Cin + ci];
            }
        wp.upload();
    }
    y.alloc((size_t)N * OH * OW * Cout, dtype, false);
    if (use_res) res.alloc(y.n, dtype, true);
    float* d_bias; CK(cudaMalloc(&d_bias, Cout * 4)); CK(cudaMemcpy(d_bias, bias.data(), Cout * 4, cudaMemcpyHostToDevice));
    float* d_stats = nullptr; const int G = 32;
    if (use_gn) { CK(cudaMalloc(&d_stats, N * G * 2 * 4)); CK(cudaMemset(d_stats, 0, N * G * 2 * 4)); }
    Buf planes;
    const void* xin = x.d;
    if (mode == 1) {
        planes.alloc(x.n, dtype, false);
        for (int pr = 0; pr < 2; ++pr) for (int pc = 0; pc < 2; ++pc) for (int n = 0; n < N; ++n)
            for (int i = 0; i < OH; ++i) for (int j = 0; j < OW; ++j) for (int c = 0; c < Cin; ++c)
                planes.h[(((((size_t)(pr * 2 + pc) * N + n) * OH + i) * OW) + j) * Cin + c] = x.h[(((size_t)n * H + 2 * i + pr) * W + 2 * j + pc) * Cin + c];
        planes.upload();
        xin = planes.d;
    }
    const int kH = mode == 1 ? OH : H, kW = mode == 1 ? OW : W;
    B2(b2_conv2d_nhwc(mode, xin, N, kH, kW, Cin, Cin, wp.d, d_bias, Cout, y.d, Cout, act, use_res ? res.d : nullptr, Cout,
                      d_stats, use_gn ? G : 0, 0, dtype, nullptr));
    CK(cudaDeviceSynchronize());
    y.download();
    // reference
    double maxerr = 0; std::vector<double> s1(N * G, 0.0), s2(N * G, 0.0);
    const int cpg = Cout / G;
    for (int n = 0; n < N; ++n) for (int oh = 0; oh < OH; ++oh) for (int ow = 0; ow < OW; ++ow) for (int co = 0; co < Cout; ++co) {
        double acc = bias[co];
        for (int kh = 0; kh < KS; ++kh) for (int kw = 0; kw < KS; ++kw) {
            int ih, iw;
            if (mode == 0) { ih = oh - 1 + kh; iw = ow - 1 + kw; }
            else if (mode == 1) { ih = 2 * oh - 1 + kh; iw = 2 * ow - 1 + kw; }
            else { int th = oh + 1 - kh, tw = ow + 1 - kw; if (th % 2 || tw % 2 || th < 0 || tw < 0) continue; ih = th / 2; iw = tw / 2; }
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            const float* xp = &x.h[(((size_t)n * H + ih) * W + iw) * Cin];
            const float* wq = &w[(((size_t)co * KS + kh) * KS + kw) * Cin];
            for (int ci = 0; ci < Cin; ++ci) acc += (double)xp[ci] * wq[ci];
        }
        if (act) acc = swish(acc);
        if (use_gn) { s1[n * G + co / cpg] += acc; s2[n * G + co / cpg] += acc * acc; }
        size_t oi = (((size_t)n * OH + oh) * OW + ow) * Cout + co;
        if (use_res) acc += res.h[oi];
        double e = fabs(acc - y.h[oi]) / (1.0 + fabs(acc));
        if (e > maxerr) maxerr = e;
    }
    char name[256];
    snprintf(name, sizeof(name), "conv mode%d %s N%d H%d W%d Cin%d Cout%d res%d gn%d act%d", mode, dtype ? "tf32" : "bf16", N, H, W, Cin, Cout, use_res, use_gn, act);
    report(name, maxerr, dtype == 0 ? 1e-2 : 2e-3);
    if (use_gn) {
        std::vector<float> st(N * G * 2); CK(cudaMemcpy(st.data(), d_stats, st.size() * 4, cudaMemcpyDeviceToHost));
        double me = 0;
        for (int i = 0; i < N * G; ++i) {
            me = fmax(me, fabs(st[2 * i] - s1[i]) / (1 + fabs(s1[i])));
            me = fmax(me, fabs(st[2 * i + 1] - s2[i]) / (1 + fabs(s2[i])));
        }
        report("   gn partial sums", me, 1e-3);
        cudaFree(d_stats);
    }
    cudaFree(d_bias); x.free_(); wp.free_(); y.free_(); res.free_(); planes.free_();
}

static void test_gemm(int dtype, int M, int Nc, int K, int b1, int b2, bool use_bias, float alpha, int act, bool use_res, int out_fp32) {
    Buf A, B, C, R;
    A.alloc((size_t)b2 * b1 * M * K, dtype, true, 2.f);
    B.alloc((size_t)b2 * b1 * Nc * K, dtype, true, 0.5f);
    const bool batched = b1 > 1 || b2 > 1;
    const int cdt = (dtype == 1 || out_fp32) ? 1 : 0;
    C.alloc((size_t)b2 * b1 * M * Nc, cdt, false);
    if (use_res) R.alloc(C.n, cdt, true);
    std::vector<float> bias(Nc); for (auto& v : bias) v = frand();
    float* d_bias; CK(cudaMalloc(&d_bias, Nc * 4)); CK(cudaMemcpy(d_bias, bias.data(), Nc * 4, cudaMemcpyHostToDevice));
    B2(b2_gemm_nt(A.d, K, (long long)M * K, (long long)b1 * M * K, B.d, K, (long long)Nc * K, (long long)b1 * Nc * K, C.d, Nc,
                  (long long)M * Nc, (long long)b1 * M * Nc, M, Nc, K, b1, b2, use_bias ? d_bias : nullptr, alpha, act,
                  use_res ? R.d : nullptr, Nc, out_fp32, dtype, nullptr));
    CK(cudaDeviceSynchronize());
    C.download();
    double maxerr = 0;
    for (int b = 0; b < b1 * b2; ++b) for (int m = 0; m < M; ++m) for (int n = 0; n < Nc; ++n) {
        double acc = 0;
        const float* ap = &A.h[((size_t)b * M + m) * K];
        const float* bp = &B.h[((size_t)(batched ? b : 0) * Nc + n) * K];
        for (int k = 0; k < K; ++k) acc += (double)ap[k] * bp[k];
        acc *= alpha;
        if (use_bias) acc += bias[n];
        if (act) acc = swish(acc);
        size_t oi = ((size_t)b * M + m) * Nc + n;
        if (use_res) acc += R.h[oi];
        double e = fabs(acc - C.h[oi]) / (1.0 + fabs(acc));
        if (e > maxerr) maxerr = e;
    }
    char name[256];
    snprintf(name, sizeof(name), "gemm %s M%d N%d K%d b%dx%d bias%d alpha%.2f act%d res%d f32out%d", dtype ? "tf32" : "bf16", M, Nc, K, b1, b2, use_bias, alpha, act, use_res, out_fp32);
    report(name, maxerr, cdt == 0 ? 1e-2 : 2e-3);
    cudaFree(d_bias); A.free_(); B.free_(); C.free_(); R.free_();
}

static void perf_conv(int dtype, int N, int H, int W, int C, int iters) {
    const size_t eb = dtype == 0 ? 2 : 4;
    void *x, *w, *y; float* bias;
    CK(cudaMalloc(&x, (size_t)N * H * W * C * eb)); CK(cudaMalloc(&y, (size_t)N * H * W * C * eb));
    CK(cudaMalloc(&w, (size_t)C * 9 * C * eb)); CK(cudaMalloc(&bias, C * 4));
    CK(cudaMemset(x, 0, (size_t)N * H * W * C * eb)); CK(cudaMemset(w, 0, (size_t)C * 9 * C * eb)); CK(cudaMemset(bias, 0, C * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) B2(b2_conv2d_nhwc(0, x, N, H, W, C, C, w, bias, C, y, C, 1, nullptr, 0, nullptr, 0, 0, dtype, nullptr));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) B2(b2_conv2d_nhwc(0, x, N, H, W, C, C, w, bias, C, y, C, 1, nullptr, 0, nullptr, 0, 0, dtype, nullptr));
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    double fl = 2.0 * N * H * W * (double)C * C * 9;
    printf("perf conv3x3 %s N%d %dx%d C%d: %.3f ms  %.1f TFLOP/s\n", dtype ? "tf32" : "bf16", N, H, W, C, ms, fl / ms * 1e-9);
    cudaFree(x); cudaFree(y); cudaFree(w); cudaFree(bias);
}

// weight gradient: reference in double from rounded operands
static void test_wgrad(int mode, int dtype, int N, int H, int W, int Cin, int Cout) {
    // (H, W): forward INPUT dims of the conv (mode 1: even; mode 2: input of the transposed conv)
    const int OH = mode == 0 ? H : (mode == 1 ? H / 2 : 2 * H), OW = mode == 0 ? W : (mode == 1 ? W / 2 : 2 * W);
    const int KS = mode == 2 ? 4 : 3;
    Buf x, dz, planes;
    x.alloc((size_t)N * H * W * Cin, dtype, true, 2.f);
    dz.alloc((size_t)N * OH * OW * Cout, dtype, true, 1.f);
    const void* xin = x.d;
    if (mode == 1) {
        planes.alloc(x.n, dtype, false);
        for (int pr = 0; pr < 2; ++pr) for (int pc = 0; pc < 2; ++pc) for (int n = 0; n < N; ++n)
            for (int i = 0; i < OH; ++i) for (int j = 0; j < OW; ++j) for (int c = 0; c < Cin; ++c)
                planes.h[(((((size_t)(pr * 2 + pc) * N + n) * OH + i) * OW) + j) * Cin + c] = x.h[(((size_t)n * H + 2 * i + pr) * W + 2 * j + pc) * Cin + c];
        planes.upload();
        xin = planes.d;
    }
    const size_t gsz = (size_t)Cout * KS * KS * Cin;
    float* d_g; CK(cudaMalloc(&d_g, gsz * 4)); CK(cudaMemset(d_g, 0, gsz * 4));
    const int kH = mode == 1 ? OH : H, kW = mode == 1 ? OW : W;
    B2(b2_conv2d_wgrad(mode, xin, N, kH, kW, Cin, Cin, dz.d, Cout, Cout, d_g, dtype, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> g(gsz); CK(cudaMemcpy(g.data(), d_g, gsz * 4, cudaMemcpyDeviceToHost));
    // reference dW[co][kh][kw][ci]
    double maxerr = 0, maxref = 0;
    for (int co = 0; co < Cout; co += 7) for (int kh = 0; kh < KS; ++kh) for (int kw = 0; kw < KS; ++kw) for (int ci = 0; ci < Cin; ci += 5) {
        double acc = 0;
        for (int n = 0; n < N; ++n) for (int oh = 0; oh < OH; ++oh) for (int ow = 0; ow < OW; ++ow) {
            int ih, iw;
            if (mode == 0) { ih = oh - 1 + kh; iw = ow - 1 + kw; }
            else if (mode == 1) { ih = 2 * oh - 1 + kh; iw = 2 * ow - 1 + kw; }
            else { int th = oh + 1 - kh, tw = ow + 1 - kw; if (th % 2 || tw % 2 || th < 0 || tw < 0) continue; ih = th / 2; iw = tw / 2; }
            if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
            acc += (double)dz.h[(((size_t)n * OH + oh) * OW + ow) * Cout + co] * x.h[(((size_t)n * H + ih) * W + iw) * Cin + ci];
        }
        size_t gi;
        if (mode != 2) gi = (((size_t)co * 9) + kh * 3 + kw) * Cin + ci;
        else {
            // packed [g=(a,b)][co][t=(i,j)][ci]; a=0: kh {1,3} ; a=1: kh {2,0}
            int a = (kh == 1 || kh == 3) ? 0 : 1, i = (kh == 1 || kh == 2) ? 0 : 1;
            int b = (kw == 1 || kw == 3) ? 0 : 1, j = (kw == 1 || kw == 2) ? 0 : 1;
            gi = ((((size_t)(a * 2 + b) * Cout + co) * 4) + i * 2 + j) * Cin + ci;
        }
        maxerr = fmax(maxerr, fabs(acc - g[gi]));
        maxref = fmax(maxref, fabs(acc));
    }
    char name[256];
    snprintf(name, sizeof(name), "wgrad mode%d %s N%d H%d W%d Cin%d Cout%d (ref max %.2f)", mode, dtype ? "tf32" : "bf16", N, H, W, Cin, Cout, maxref);
    report(name, maxerr / (maxref + 1e-9), 2e-3);
    cudaFree(d_g); x.free_(); dz.free_(); planes.free_();
}

static void test_gemm_tn(int dtype, int M, int Nc, int K, int b1, int b2, int out_mode, float alpha) {
    Buf A, B, C;
    A.alloc((size_t)b2 * b1 * K * M, dtype, true, 1.f);
    B.alloc((size_t)b2 * b1 * K * Nc, dtype, true, 1.f);
    const int cdt = out_mode == 0 ? 1 : dtype;
    C.alloc((size_t)b2 * b1 * M * Nc, cdt, false);
    B2(b2_gemm_tn(A.d, M, (long long)K * M, (long long)b1 * K * M, B.d, Nc, (long long)K * Nc, (long long)b1 * K * Nc, C.d, Nc,
                  (long long)M * Nc, (long long)b1 * M * Nc, M, Nc, K, b1, b2, alpha, out_mode, dtype, nullptr));
    CK(cudaDeviceSynchronize());
    C.download();
    double maxerr = 0;
    for (int b = 0; b < b1 * b2; ++b) for (int m = 0; m < M; ++m) for (int n = 0; n < Nc; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += (double)A.h[((size_t)b * K + k) * M + m] * B.h[((size_t)b * K + k) * Nc + n];
        acc *= alpha;
        maxerr = fmax(maxerr, fabs(acc - C.h[((size_t)b * M + m) * Nc + n]) / (1.0 + fabs(acc)));
    }
    if (maxerr > 0.1) {
        printf("  debug gemm_tn: first row got/want:");
        for (int n = 0; n < 8; ++n) { double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A.h[(size_t)k * M + 0] * B.h[(size_t)k * Nc + n]; printf(" %.3f/%.3f", C.h[n], acc * alpha); }
        printf("\n  row 33:");
        for (int n = 0; n < 8; ++n) { double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A.h[(size_t)k * M + 33] * B.h[(size_t)k * Nc + n]; printf(" %.3f/%.3f", C.h[(size_t)33 * Nc + n], acc * alpha); }
        printf("\n");
    }
    char name[256];
    snprintf(name, sizeof(name), "gemm_tn %s M%d N%d K%d b%dx%d mode%d", dtype ? "tf32" : "bf16", M, Nc, K, b1, b2, out_mode);
    report(name, maxerr, (cdt == 0) ? 1e-2 : 2e-3);
    A.free_(); B.free_(); C.free_();
}

static void perf_wgrad(int dtype, int N, int H, int W, int C, int iters) {
    const size_t eb = dtype == 0 ? 2 : 4;
    void *x, *dz; float* g;
    CK(cudaMalloc(&x, (size_t)N * H * W * C * eb)); CK(cudaMalloc(&dz, (size_t)N * H * W * C * eb));
    CK(cudaMalloc(&g, (size_t)C * 9 * C * 4));
    CK(cudaMemset(x, 0, (size_t)N * H * W * C * eb)); CK(cudaMemset(dz, 0, (size_t)N * H * W * C * eb)); CK(cudaMemset(g, 0, (size_t)C * 9 * C * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) B2(b2_conv2d_wgrad(0, x, N, H, W, C, C, dz, C, C, g, dtype, nullptr));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) B2(b2_conv2d_wgrad(0, x, N, H, W, C, C, dz, C, C, g, dtype, nullptr));
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    double fl = 2.0 * N * H * W * (double)C * C * 9;
    printf("perf wgrad3x3 %s N%d %dx%d C%d: %.3f ms  %.1f TFLOP/s\n", dtype ? "tf32" : "bf16", N, H, W, C, ms, fl / ms * 1e-9);
    cudaFree(x); cudaFree(dz); cudaFree(g);
}

int main(int argc, char** argv) {
    const bool perf = argc > 1 && std::string(argv[1]) == "perf";
    // plain GEMM first: the simplest use of the pipeline
    test_gemm(0, 128, 128, 64, 1, 1, false, 1.f, 0, false, 0);
    test_gemm(0, 128, 128, 256, 1, 1, false, 1.f, 0, false, 0);
    test_gemm(0, 300, 192, 256, 1, 1, true, 0.5f, 1, true, 0);
    test_gemm(0, 1000, 1536, 512, 1, 1, true, 1.f, 0, false, 0);
    test_gemm(0, 256, 256, 512, 3, 2, false, 0.04f, 0, false, 1);
    test_gemm(0, 64, 64, 128, 5, 1, false, 1.f, 0, false, 1);
    test_gemm(1, 300, 192, 256, 1, 1, true, 0.5f, 1, true, 0);
    test_gemm(1, 256, 256, 512, 3, 2, false, 0.04f, 0, false, 0);
    // convolutions
    test_conv(0, 0, 2, 16, 16, 64, 128, false, true, 1);
    test_conv(0, 0, 1, 64, 64, 128, 128, true, true, 1);
    test_conv(0, 0, 5, 4, 4, 128, 256, false, true, 1);
    test_conv(0, 0, 3, 2, 2, 128, 128, false, true, 0);
    test_conv(0, 0, 2, 32, 32, 256, 256, true, true, 1);
    test_conv(0, 1, 2, 16, 16, 64, 128, true, true, 1);
    test_conv(1, 0, 2, 16, 16, 128, 256, false, false, 1);
    test_conv(1, 0, 3, 4, 4, 128, 128, false, false, 1);
    test_conv(1, 1, 2, 16, 16, 64, 128, false, false, 1);
    test_conv(2, 0, 2, 8, 8, 128, 64, false, false, 1);
    test_conv(2, 0, 3, 2, 2, 256, 128, false, false, 1);
    test_conv(2, 1, 2, 8, 8, 64, 64, false, false, 1);
    // TN GEMM / weight gradients
    test_gemm_tn(0, 128, 128, 64, 1, 1, 0, 1.f);
    test_gemm_tn(0, 128, 256, 512, 1, 1, 0, 1.f);
    test_gemm_tn(0, 200, 192, 300, 1, 1, 0, 0.5f);
    test_gemm_tn(0, 64, 512, 64, 2, 3, 1, 1.f);
    test_gemm_tn(0, 16, 64, 16, 2, 2, 1, 1.f);
    test_gemm_tn(0, 8, 64, 4, 1, 3, 1, 0.25f);
    test_gemm_tn(1, 128, 128, 256, 1, 1, 0, 1.f);
    test_gemm_tn(1, 64, 96, 40, 2, 2, 1, 1.f);
    test_wgrad(0, 0, 2, 16, 16, 64, 128);
    test_wgrad(0, 0, 1, 64, 64, 128, 128);
    test_wgrad(0, 0, 3, 2, 2, 128, 64);
    test_wgrad(0, 0, 5, 4, 4, 256, 256);
    test_wgrad(0, 1, 2, 8, 8, 64, 128);
    test_wgrad(1, 0, 2, 16, 16, 128, 128);
    test_wgrad(1, 0, 3, 4, 4, 64, 128);
    test_wgrad(2, 0, 2, 8, 8, 128, 64);
    test_wgrad(2, 0, 3, 2, 2, 128, 128);
    test_wgrad(2, 1, 2, 4, 4, 64, 64);
    printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    if (perf) {
        perf_conv(0, 256, 16, 16, 1024, 10);
        perf_conv(0, 256, 32, 32, 512, 10);
        perf_conv(0, 256, 64, 64, 128, 10);
        perf_conv(0, 8, 16, 16, 1024, 10);
        perf_conv(1, 64, 16, 16, 1024, 5);
        perf_wgrad(0, 8, 16, 16, 1024, 10);
        perf_wgrad(0, 32, 16, 16, 1024, 10);
        perf_wgrad(0, 8, 64, 64, 128, 10);
        perf_wgrad(0, 32, 32, 32, 512, 10);
    }
    return g_fail ? 1 : 0;
}
