// Host launchers (C-ABI) for the tcgen05 implicit-GEMM kernel: 3x3 conv (stride 1 / stride 2 on parity
// planes), 4x4 stride-2 transposed conv (four output-parity GEMMs), and plain / batched NT GEMM.
#include "host_util.h"
#include "igemm.h"
#include "sdm_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace b2 {
int launch_igemm_nt(int dtype, const CUtensorMap& a, const CUtensorMap& b, const IgemmParams& p, int block_n, cudaStream_t st);
int launch_gn_stats(const void* y, long long ldy, float* stats, int N, int HW, int C, int groups, int pre_swish, int dtype,
                    cudaStream_t st, bool deterministic);
}
using namespace b2;

// ---- tiling: N tile width and split-K factor -------------------------------------------------------------------
// Workspace for split-K partial sums, registered once by the host side (b2_set_workspace): [1024 int counters | fp32].
// One workspace PER DEVICE (keyed by the current device of the calling thread), each split in two halves so that the NT
// kernel (forward / data gradients, main stream) and the TN kernel (weight gradients, possibly on the side stream of
// SDM_B200_OVERLAP_WGRAD) never share partial-sum slices or counters.  Kernels that use one half are still ordered among
// themselves by their stream.
struct Workspace { int* counters; float* ws; long long floats; };
static constexpr int kMaxDevices = 64;
static Workspace g_wtab[kMaxDevices][2];

static const Workspace& workspace(int half) {
    static const Workspace none = {nullptr, nullptr, 0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return none;
    return g_wtab[dev][half];
}

extern "C" int b2_set_deterministic(int on) { set_deterministic_mode(on); return 0; }

extern "C" int b2_set_workspace(void* ws, long long bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return set_error("b2_set_workspace: no current device");
    if (!ws || bytes < (2 << 20)) { g_wtab[dev][0] = g_wtab[dev][1] = Workspace{nullptr, nullptr, 0}; return 0; }
    if ((uintptr_t)ws % 256) return set_error("b2_set_workspace: pointer must be 256-byte aligned");
    const long long half = (bytes / 2) & ~255LL;
    for (int h = 0; h < 2; ++h) {
        char* base = reinterpret_cast<char*>(ws) + h * half;
        g_wtab[dev][h].counters = reinterpret_cast<int*>(base);                      // [1024] int, zero on entry, self-resetting
        g_wtab[dev][h].ws = reinterpret_cast<float*>(base + 4096);
        g_wtab[dev][h].floats = (half - 4096) / 4;
    }
    return 0;
}

// Picks the N tile (256 / 128 / 64 accumulator columns) and, for layers whose few output tiles cannot occupy the 148
// SMs (deep U-Net levels: 2x2 ... 8x8 pixels per image), a split of the K loop, by minimising a small cost model
// calibrated on B200 with tools/sweep_tiling.py (profiles/r01d_tiling_sweep_*.log):
//   T = launch/pipeline latency + waves * k_iters/splits * t_iter(bn, machine fill) [+ split hand-off + per-slice read-back]
// One K step (64 channels of one tap) costs 0.18 / 0.30 / 0.33 us on a lightly loaded machine (TMA-latency bound:
// 8 / 6 / 4 pipeline stages) and 0.27 / 0.33 / 0.41 us with all SMs streaming (L2 bound).  The split hand-off is a
// store + fence + counter round trip (~5 us) and the last split reads every partial slice back (cost ~ slice size),
// so splitting only pays with narrow tiles.
static void pick_tiling(int cout, long long m_tiles, int groups, int k_iters, bool allow_split, int* bn_out, int* splits_out) {
    const int sms = device_sm_count();
    const Workspace& wsp = workspace(0);
    float* const g_ws = wsp.ws;
    const long long g_ws_floats = wsp.floats;
    const int cands[3] = {256, 128, 64};
    const double t_light[3] = {0.33, 0.30, 0.18}, t_full[3] = {0.41, 0.33, 0.27}, slice_us[3] = {9.0, 1.5, 0.6};
    const int split_cands[8] = {1, 2, 3, 4, 6, 8, 12, 16};
    double best = 1e30;
    int best_bn = 64, best_s = 1;
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        if (cout < bn && i < 2 && cout <= cands[i + 1]) continue;           // a narrower tile already covers all columns
        const long long tiles = m_tiles * ((cout + bn - 1) / bn) * groups;
        for (int j = 0; j < 8; ++j) {
            const int sp = split_cands[j];
            if (sp > 1) {
                if (!allow_split || !g_ws || bn == 256 || k_iters / sp < 4 || tiles > 1024) break;
                if (tiles * sp * 128LL * bn > g_ws_floats) break;
                if (tiles * (sp - 1) >= sms) break;                              // the machine is full already
            }
            const long long items = tiles * sp;
            const long long waves = (items + sms - 1) / sms;
            const double fill = items >= sms ? 1.0 : (double)items / sms;
            const double t_iter = t_light[i] + (t_full[i] - t_light[i]) * fill;
            double t = 12.0 + (double)waves * ((double)k_iters / sp) * t_iter;
            if (waves > 1) t += (double)(waves - 1) * 1.5;                       // per-tile epilogue tail that is not hidden
            if (sp > 1) t += 5.0 + sp * slice_us[i];
            if (t < best) { best = t; best_bn = bn; best_s = sp; }
        }
    }
    *bn_out = best_bn;
    *splits_out = best_s;
    // tuning hook (tools/sweep_tiling.py): SDM_B200_FORCE_TILING="<bn>,<splits>" overrides the choice where it is legal
    if (const char* force = getenv("SDM_B200_FORCE_TILING")) {
        int fbn = 0, fs = 0;
        if (sscanf(force, "%d,%d", &fbn, &fs) == 2 && (fbn == 64 || fbn == 128 || fbn == 256) && fs >= 1) {
            const long long tiles = m_tiles * ((cout + fbn - 1) / fbn) * groups;
            const bool split_ok = fs == 1 || (allow_split && g_ws && k_iters / fs >= 1 && tiles <= 1024 && tiles * fs * 128LL * fbn <= g_ws_floats);
            if (split_ok) { *bn_out = fbn; *splits_out = fs; }
        }
    }
}

// Cluster size for TMA multicast of the weight tile (igemm_nt.cu, CL): only worth it when the layer has several waves of
// tiles (the main loop is then bound by L2 -> SM operand traffic); bf16 kernels only.  SDM_B200_CLUSTER=1|2|4 overrides.
static int pick_cluster(int dtype, int bn, int splits, long long tiles, long long m_tiles) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("SDM_B200_CLUSTER");
        forced = e ? atoi(e) : 0;
    }
    if (dtype != 0 || splits != 1 || bn < 128) return 1;
    const int sms = device_sm_count();
    int cl = forced > 0 ? forced : 2;
    if (cl != 1 && cl != 2 && cl != 4) cl = 2;
    if (cl > 1 && (tiles < (long long)sms || m_tiles < 8 * cl)) return 1;
    return cl;
}

// Spatial box of <=128 output pixels: full rows first, then rows, then images.
static void pick_box(int W, int H, int N, int* wb, int* hb, int* nb, int rows = 128) {
    *wb = W < rows ? W : rows;
    int rem = rows / *wb;
    *hb = H < rem ? H : rem;
    rem /= *hb;
    *nb = N < rem ? N : rem;
    if (*nb < 1) *nb = 1;
}

struct ColSum { const void* z; long long ldz; float* s1; float* s2; };
static int conv2d_impl(int mode, const void* x, int N, int H, int W, int Cin, long long ldx,
                       const void* wpacked, const float* bias, int Cout, void* y, long long ldy, int act,
                       const void* residual, long long ldr, float* gn_stats, int gn_groups, int out_mode,
                       int dtype, void* stream, const ColSum* cs, void* y2 = nullptr, long long ldy2 = 0);

extern "C" int b2_conv2d_nhwc(int mode, const void* x, int N, int H, int W, int Cin, long long ldx,
                              const void* wpacked, const float* bias, int Cout, void* y, long long ldy, int act,
                              const void* residual, long long ldr, float* gn_stats, int gn_groups, int out_mode,
                              int dtype, void* stream) {
    return conv2d_impl(mode, x, N, H, W, Cin, ldx, wpacked, bias, Cout, y, ldy, act, residual, ldr, gn_stats, gn_groups, out_mode,
                       dtype, stream, nullptr);
}

// b2_conv2d_nhwc (act 0) with a second output y2 = Swish(y): training forward of the convs without normalisation.
extern "C" int b2_conv2d_nhwc_dual(int mode, const void* x, int N, int H, int W, int Cin, long long ldx,
                                   const void* wpacked, const float* bias, int Cout, void* y, long long ldy,
                                   void* y_act, long long ldy_act, int dtype, void* stream) {
    if (!y_act) return set_error("b2_conv2d_nhwc_dual: y_act is NULL");
    if (mode < 0 || mode > 2) return set_error("b2_conv2d_nhwc_dual: forward modes 0..2 only");
    return conv2d_impl(mode, x, N, H, W, Cin, ldx, wpacked, bias, Cout, y, ldy, 0, nullptr, 0, nullptr, 0, 0, dtype, stream, nullptr,
                       y_act, ldy_act);
}

// b2_conv2d_nhwc (mode 0, bf16) whose epilogue ALSO accumulates, per (image, output channel), cs_s1 += sum_p y and
// cs_s2 += sum_p y * swish(cs_z[p]) -- the pass-1 sums of the AdaGN backward that consumes y as its `dout` (cs_z = that layer's
// pre-activation, [N][H][W][Cout] with per-pixel stride cs_ldz; cs_s1 / cs_s2: [N][Cout] fp32, zeroed by the caller).
extern "C" int b2_conv2d_nhwc_colsum(int mode, const void* x, int N, int H, int W, int Cin, long long ldx,
                                     const void* wpacked, const float* bias, int Cout, void* y, long long ldy, int act,
                                     const void* residual, long long ldr, float* gn_stats, int gn_groups, int out_mode,
                                     int dtype, void* stream, const void* cs_z, long long cs_ldz, float* cs_s1, float* cs_s2) {
    if ((mode != 0 && mode != 5) || dtype != 0 || out_mode != 0) return set_error("b2_conv2d_nhwc_colsum: plain bf16 3x3 stride-1 convolutions only");
    if (!cs_z || !cs_s1 || !cs_s2 || Cout % 32 || (cs_ldz % 8) || ((uintptr_t)cs_z % 16)) return set_error("b2_conv2d_nhwc_colsum: bad column-sum arguments");
    ColSum cs = {cs_z, cs_ldz, cs_s1, cs_s2};
    return conv2d_impl(mode, x, N, H, W, Cin, ldx, wpacked, bias, Cout, y, ldy, act, residual, ldr, gn_stats, gn_groups, out_mode,
                       dtype, stream, &cs);
}

static int conv2d_impl(int mode, const void* x, int N, int H, int W, int Cin, long long ldx,
                       const void* wpacked, const float* bias, int Cout, void* y, long long ldy, int act,
                       const void* residual, long long ldr, float* gn_stats, int gn_groups, int out_mode,
                       int dtype, void* stream, const ColSum* cs, void* y2, long long ldy2) {
    const int eb = dtype == 0 ? 2 : 4;
    const int bk = 128 / eb;
    if (Cin % bk != 0) return set_error("b2_conv2d_nhwc: Cin=%d must be a multiple of %d", Cin, bk);
    if (out_mode != 0 && (mode != 0 || residual)) return set_error("b2_conv2d_nhwc: NCHW fp32 output only for the plain 3x3 conv");
    if (mode < 0 || mode > 5) return set_error("b2_conv2d_nhwc: bad mode %d", mode);
    // mode 5 = data gradient of the 3x3/s1 conv straight from the FORWARD weights (igemm.h: b_mn): x = dz [N][H][W][Cin = Cout_f],
    // wpacked = the bf16 forward weights [Cout_f][9][Cin_f], Cout = Cin_f.  Everything else is mode 0.
    const bool b_mn = mode == 5;
    if (b_mn) {
        if (dtype != 0 || out_mode != 0 || Cout % 64 || Cin % 64) return set_error("b2_conv2d_nhwc mode 5: bf16, whole 64-channel slabs only");
        mode = 0;
    }
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.b_mn = b_mn ? 1 : 0;
    p.W = W; p.H = H; p.N = N;
    pick_box(W, H, N, &p.wb, &p.hb, &p.nb);
    p.tiles_w = (W + p.wb - 1) / p.wb;
    p.tiles_h = (H + p.hb - 1) / p.hb;
    p.tiles_n = (N + p.nb - 1) / p.nb;
    p.kb_per_tap = Cin / bk;
    p.Cout = Cout;
    p.out = y;
    p.bias = bias;
    p.alpha = 1.0f;
    p.act = act;
    p.residual = residual;
    p.gn_stats = gn_stats;
    p.cpg = gn_groups > 0 ? Cout / gn_groups : 0;
    // the epilogue fuses the statistics for power-of-two group widths >= 4 (every width of the reference's configs with
    // C >= 128); narrower nets get them from a separate streaming pass over the conv output
    const bool det = deterministic_mode();      // no fp atomics: statistics from the fixed-order pass (host_util.h)
    const bool separate_stats = gn_stats && (det || p.cpg < 4 || (p.cpg & (p.cpg - 1)) != 0);
    if (separate_stats) {
        if (out_mode != 0 || mode != 0) return set_error("b2_conv2d_nhwc: unfused GroupNorm statistics only for the plain 3x3 conv");
        p.gn_stats = nullptr;
    }
    int a_images = N;
    if (mode == 0) {          // 3x3, stride 1, pad 1
        p.groups = 1; p.taps = 9;
        for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
            p.tap_dh[kh * 3 + kw] = kh - 1; p.tap_dw[kh * 3 + kw] = kw - 1; p.tap_dn[kh * 3 + kw] = 0;
        }
        p.oN = (long long)H * W * ldy; p.oH = (long long)W * ldy; p.oW = ldy;
        p.rN = (long long)H * W * ldr; p.rH = (long long)W * ldr; p.rW = ldr;
    } else if (mode == 1) {   // 3x3, stride 2, pad 1; x = parity planes [2][2][N][H][W][Cin], (H, W) = output size
        p.groups = 1; p.taps = 9;
        for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
            const int pr = (kh != 1), pc = (kw != 1);
            p.tap_dh[kh * 3 + kw] = (kh == 0) ? -1 : 0;
            p.tap_dw[kh * 3 + kw] = (kw == 0) ? -1 : 0;
            p.tap_dn[kh * 3 + kw] = (pr * 2 + pc) * N;
        }
        a_images = 4 * N;
        p.oN = (long long)H * W * ldy; p.oH = (long long)W * ldy; p.oW = ldy;
        p.rN = (long long)H * W * ldr; p.rH = (long long)W * ldr; p.rW = ldr;
    } else if (mode == 4) {   // data gradient of the transposed 4x4/s2 conv = 4x4/s2 conv over dz: 16 taps on dz's parity planes
        p.groups = 1; p.taps = 16;
        for (int kh = 0; kh < 4; ++kh) for (int kw = 0; kw < 4; ++kw) {
            const int t = kh * 4 + kw;
            const int pr = (kh == 0 || kh == 2), pc = (kw == 0 || kw == 2);
            p.tap_dh[t] = kh == 0 ? -1 : (kh == 3 ? 1 : 0);
            p.tap_dw[t] = kw == 0 ? -1 : (kw == 3 ? 1 : 0);
            p.tap_dn[t] = (pr * 2 + pc) * N;
        }
        a_images = 4 * N;
        p.oN = (long long)H * W * ldy; p.oH = (long long)W * ldy; p.oW = ldy;
        p.rN = (long long)H * W * ldr; p.rH = (long long)W * ldr; p.rW = ldr;
    } else {                  // mode 2: transposed 4x4, stride 2, pad 1; output 2H x 2W; group = output parity (a, b)
                              // mode 3: data gradient of the 3x3/s2 conv (same addressing; 1-2 real taps per dim, zero-padded weights)
        p.groups = 4; p.taps = 4;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
            const int g = a * 2 + b;
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                const int t = g * 4 + i * 2 + j;
                if (mode == 2) {
                    p.tap_dh[t] = (i == 0) ? 0 : (a == 0 ? -1 : 1);
                    p.tap_dw[t] = (j == 0) ? 0 : (b == 0 ? -1 : 1);
                } else {
                    p.tap_dh[t] = (i == 1 && a == 1) ? 1 : 0;
                    p.tap_dw[t] = (j == 1 && b == 1) ? 1 : 0;
                }
                p.tap_dn[t] = 0;
            }
            p.goff[g] = ((long long)a * (2 * W) + b) * ldy;
        }
        p.oN = (long long)4 * H * W * ldy; p.oH = (long long)2 * (2 * W) * ldy; p.oW = 2 * ldy;
        p.rN = (long long)4 * H * W * ldr; p.rH = (long long)2 * (2 * W) * ldr; p.rW = 2 * ldr;
        if (residual) return set_error("b2_conv2d_nhwc: residual not supported for the parity-decomposed modes");
    }
    p.oC = 1;
    if (out_mode == 1) {      // final layer: fp32 NCHW straight from the epilogue (ldy ignored)
        p.oN = (long long)Cout * H * W; p.oH = W; p.oW = 1; p.oC = (long long)H * W;
        p.out_fp32 = 1;
    }
    {
        const int ov = (dtype == 1 || p.out_fp32) ? 4 : 8;    // elements per 16 bytes
        const int oeb = 16 / ov;
        bool ok = (p.oC == 1) && (ldy % ov == 0) && ((uintptr_t)y % 16 == 0);
        for (int g = 0; g < p.groups; ++g) ok = ok && (p.goff[g] % ov == 0);
        if (residual) ok = ok && (ldr % ov == 0) && ((uintptr_t)residual % 16 == 0);
        (void)oeb;
        p.vec_ok = ok ? 1 : 0;
    }
    if (y2) {                 // second output = Swish(y), NHWC with its own per-pixel stride (igemm.h: out2)
        if (out_mode != 0 || act != 0 || gn_stats || residual) return set_error("b2_conv2d_nhwc_dual: plain conv + bias only (act 0, no statistics / residual)");
        const int ov = dtype == 1 ? 4 : 8;
        const bool parity = (p.groups == 4);        // modes 2 / 3: output at 2H x 2W, group = output parity
        p.out2 = y2;
        if (parity) {
            p.o2N = (long long)4 * H * W * ldy2; p.o2H = (long long)2 * (2 * W) * ldy2; p.o2W = 2 * ldy2;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) p.goff2[a * 2 + b] = ((long long)a * (2 * W) + b) * ldy2;
        } else {
            p.o2N = (long long)H * W * ldy2; p.o2H = (long long)W * ldy2; p.o2W = ldy2;
        }
        bool ok2 = (ldy2 % ov == 0) && ((uintptr_t)y2 % 16 == 0);
        p.vec2_ok = ok2 ? 1 : 0;
    }
    if (cs) {
        p.cs_z = cs->z; p.cs_s1 = cs->s1; p.cs_s2 = cs->s2;
        p.cs_zN = (long long)H * W * cs->ldz; p.cs_zH = (long long)W * cs->ldz; p.cs_zW = cs->ldz;
    }
    long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    int bn, splits;
    pick_tiling(Cout, m_tiles, p.groups, p.taps * p.kb_per_tap, !det, &bn, &splits);
    // ---- swapped-operand mode (igemm.h: swap_ab) for 3x3 stride-1 convs with <= 128 output channels: the 128x256 MMA shape.
    // OFF by default (SDM_B200_SWAP_AB=1 / b2_set_option("swap_ab", 1) enables it).  MEASURED on the 128-channel layers
    // (profiles/r02n_conv128_variants.md; TFLOP/s forward, batch 256 @64x64 | batch 32 @128x128): per-tap 884 | 874, halo
    // 965 | 976, swapped 848 | 799 -- the wider MMA shape does not help: these layers are bound by operand delivery (18 K steps
    // per tile, every CTA streaming the same 295 KB weight matrix), which only the halo tile reduces.
    bool swapped = false;
    {
        const int swap_env = option("swap_ab", 0);
        if (swap_env && !cs && !b_mn && !y2 && mode == 0 && dtype == 0 && out_mode == 0 && Cout <= 128 && Cout >= 64 && p.cpg != 1 && !separate_stats && W <= 256) {
            int wb2, hb2, nb2;
            pick_box(W, H, N, &wb2, &hb2, &nb2, 256);
            const long long tiles2 = (long long)((W + wb2 - 1) / wb2) * ((H + hb2 - 1) / hb2) * ((N + nb2 - 1) / nb2);
            // enough 256-pixel tiles to fill the machine without split-K; whole 32-pixel chunks per image; 32-channel quarters
            if (tiles2 >= device_sm_count() && (wb2 * hb2) % 32 == 0 && Cout % 32 == 0) {
                swapped = true;
                p.swap_ab = 1;
                p.wb = wb2; p.hb = hb2; p.nb = nb2;
                p.tiles_w = (W + wb2 - 1) / wb2; p.tiles_h = (H + hb2 - 1) / hb2; p.tiles_n = (N + nb2 - 1) / nb2;
                m_tiles = tiles2;
                bn = 256; splits = 1;
            }
        }
    }
    // ---- halo mode (igemm.h): stride-1 3x3 convs with 128 / 256 input channels, whose 18 / 36 K steps per tile are too few to
    // amortise re-fetching the A box once per tap.  SDM_B200_HALO=0 keeps the per-tap loads.
    uint32_t a_box[4] = {(uint32_t)bk, (uint32_t)p.wb, (uint32_t)p.hb, (uint32_t)p.nb};
    {
        const int halo_env = option("halo", 1);
        // MEASURED (profiles/r02g_halo_ab.log): with ONE 128-row tile per weights stage the halo kernel is slower than per-tap
        // loads (C128 64x64 batch 256: 854 -> 700 TFLOP/s) although it moves 2x fewer bytes -- the main loop is bound by TMA
        // latency x ring depth in STAGES, not by bytes.  What pays is doing twice the MMA work per weights stage: two 128-row
        // sub-tiles per item (256 flattened positions), which needs 2 x 2 x BLOCK_N TMEM columns, i.e. BLOCK_N = 128 = Cout.
        const bool shape_ok = mode == 0 && dtype == 0 && out_mode == 0 && (Cin == 128 || Cin == 256) && Cout == 128 &&
                              W >= 16 && W + 1 <= 256;
        if (halo_env && shape_ok && !swapped) {
            const int hbn = 128, msub = 2;
            const int b_bytes = hbn * 128;
            const int max_b = 6;
            int halo, P = W + 1, BW, R;
            long long tw, th;
            int hb_dec = 1;
            if (W == 128 && H % 2 == 0) { halo = 2; BW = 130; R = 4; tw = 1; th = H / 2; hb_dec = 2; }       // rows h, h+1 of one image
            else { halo = 1; BW = P; R = (128 * msub + 2 * P + 1) / P + 2; tw = ((long long)H * P + 128 * msub - 1) / (128 * msub); th = 1; }
            const int a_buf = ((BW * R * 128) + 1023) & ~1023;
            const int avail = 227 * 1024 - 1024 - 256 - 2 * hbn * 4 - 1024;
            // the weights ring first (a stage is 512 MMA cycles of work against ~2000+ cycles of TMA latency), then two A boxes
            int b_stages = max_b;
            int a_slots = (avail - b_stages * b_bytes) / a_buf;
            while (a_slots < 2 && b_stages > 3) { --b_stages; a_slots = (avail - b_stages * b_bytes) / a_buf; }
            if (a_slots > 4) a_slots = 4;
            const long long h_tiles = tw * th * N;
            if (a_slots >= 2 && b_stages >= 3 && BW <= 256 && R <= 256 && h_tiles >= device_sm_count()) {
                p.halo = halo; p.halo_P = P; p.halo_BW = BW; p.halo_R = R; p.halo_msub = msub;
                p.a_buf_bytes = a_buf; p.a_slots = a_slots; p.b_stages = b_stages;
                p.halo_bar_off = a_slots * a_buf + b_stages * b_bytes;
                p.wb = halo == 1 ? 128 * msub : 128; p.hb = hb_dec; p.nb = 1;
                p.tiles_w = (int)tw; p.tiles_h = (int)th; p.tiles_n = N;
                m_tiles = h_tiles;
                bn = hbn; splits = 1;
                a_box[1] = (uint32_t)BW; a_box[2] = (uint32_t)R; a_box[3] = 1;
            }
        }
    }
    p.n_tiles = swapped ? (Cout + 127) / 128 : (Cout + bn - 1) / bn;
    p.b_mode = 0;
    p.splits = splits; p.ws = workspace(0).ws; p.ws_counters = workspace(0).counters;
    p.cluster = swapped ? 1 : pick_cluster(dtype, bn, splits, m_tiles * p.groups * p.n_tiles, m_tiles);
    if (p.halo && p.cluster > 2) p.cluster = 2;               // halo kernels are instantiated for clusters of 1 and 2
    if (b_mn && p.cluster > 2) p.cluster = 2;                 // MN-major weights: instantiated for clusters of 1 and 2
    if (b_mn && p.cluster > bn / 64) p.cluster = bn / 64;

    CUtensorMap ta, tb;
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)a_images};
        uint64_t str[3] = {(uint64_t)ldx * eb, (uint64_t)W * ldx * eb, (uint64_t)H * W * ldx * eb};
        if (make_tmap_4d(&ta, x, eb, dims, str, a_box)) return 1;
    }
    if (b_mn) {
        // (64 Cin_f of one slab | Cout_f rows | Cin_f slabs | tap) over [Cout_f][9][Cin_f]; box = 64 K rows x (bn/64)/CL slabs
        uint64_t dims[4] = {64, (uint64_t)Cin, (uint64_t)Cout / 64, 9};
        uint64_t str[3] = {(uint64_t)9 * Cout * eb, 128, (uint64_t)Cout * eb};
        uint32_t box[4] = {64, 64, (uint32_t)(bn / 64 / p.cluster), 1};
        if (make_tmap_4d(&tb, wpacked, eb, dims, str, box)) return 1;
    } else {
        const uint64_t ktot = (uint64_t)p.taps * Cin;
        uint64_t dims[4] = {ktot, (uint64_t)Cout, (uint64_t)p.groups, 1};
        uint64_t str[3] = {ktot * eb, ktot * Cout * eb, ktot * Cout * p.groups * eb};
        uint32_t box[4] = {(uint32_t)bk, (uint32_t)(swapped ? 128 : bn / p.cluster), 1, 1};      // cluster mode: each CTA fetches 1/CL of the B tile
        if (make_tmap_4d(&tb, wpacked, eb, dims, str, box)) return 1;
    }
    {   // weights of >= 1 MiB are prefetched into L2 by the kernel itself (igemm.h: pf_ptr); measured neutral on B200 (profiles/r02z4_step_ab_alignment.log), so opt-in: "l2_prefetch" 1 / SDM_B200_L2_PREFETCH=1
        const long long wbytes = (b_mn ? 9LL * Cin * Cout : (long long)p.taps * Cin * Cout * p.groups) * eb;
        if (option("l2_prefetch", 0) && wbytes >= (1 << 20) && ((uintptr_t)wpacked % 16) == 0) { p.pf_ptr = wpacked; p.pf_bytes = wbytes & ~15LL; }
    }
    if (launch_igemm_nt(dtype, ta, tb, p, bn, (cudaStream_t)stream)) return 1;
    if (separate_stats) return launch_gn_stats(y, ldy, gn_stats, N, H * W, Cout, gn_groups, act == 3 ? 1 : 0, dtype, (cudaStream_t)stream, det);
    return 0;
}

// C[b2][b1][m][n] = alpha * sum_k A[b2][b1][m][k] * B[b2][b1][n][k] (+ bias[n]) (act) (+ residual)
// Strides are in elements. b1/b2 = 1 for a plain GEMM (then B is shared: its batch strides are ignored).
extern "C" int b2_gemm_nt(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb,
                          long long b_s1, long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2,
                          int M, int Ncols, int K, int batch1, int batch2, const float* bias, float alpha, int act,
                          const void* residual, long long ldr, int out_fp32, int dtype, void* stream) {
    const int eb = dtype == 0 ? 2 : 4;
    const int bk = 128 / eb;
    if ((lda * eb) % 16 || (ldb * eb) % 16) return set_error("b2_gemm_nt: lda/ldb rows must be 16-byte multiples");
    if (residual && (batch1 != 1 || batch2 != 1)) return set_error("b2_gemm_nt: residual only for unbatched GEMM");
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.W = M; p.H = batch1; p.N = batch2;
    p.wb = 128; p.hb = 1; p.nb = 1;
    p.tiles_w = (M + 127) / 128; p.tiles_h = batch1; p.tiles_n = batch2;
    p.groups = 1; p.taps = 1; p.kb_per_tap = (K + bk - 1) / bk;   // K tail is zero-filled by TMA
    p.Cout = Ncols;
    p.out = C; p.oN = c_s2; p.oH = c_s1; p.oW = ldc; p.oC = 1;
    {
        const int ov = (dtype == 1 || out_fp32) ? 4 : 8;
        bool ok = (ldc % ov == 0) && (c_s1 % ov == 0) && (c_s2 % ov == 0) && ((uintptr_t)C % 16 == 0);
        if (residual) ok = ok && (ldr % ov == 0) && ((uintptr_t)residual % 16 == 0);
        p.vec_ok = ok ? 1 : 0;
    }
    p.residual = residual; p.rN = 0; p.rH = 0; p.rW = ldr;
    p.bias = bias; p.alpha = alpha; p.act = act; p.out_fp32 = out_fp32;
    const bool batched = (batch1 > 1 || batch2 > 1);
    p.b_mode = batched ? 1 : 0;
    int bn, splits;
    pick_tiling(Ncols, (long long)p.tiles_w * batch1 * batch2, 1, p.kb_per_tap, !batched && !deterministic_mode(), &bn, &splits);
    p.n_tiles = (Ncols + bn - 1) / bn;
    p.splits = splits; p.ws = workspace(0).ws; p.ws_counters = workspace(0).counters;
    p.cluster = batched ? 1 : pick_cluster(dtype, bn, splits, (long long)p.tiles_w * p.n_tiles, p.tiles_w);
    CUtensorMap ta, tb;
    {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)batch1, (uint64_t)batch2};
        uint64_t str[3] = {(uint64_t)lda * eb, (uint64_t)(batch1 > 1 ? a_s1 : lda * M) * eb,
                           (uint64_t)(batch2 > 1 ? a_s2 : lda * M * batch1) * eb};
        uint32_t box[4] = {(uint32_t)bk, 128, 1, 1};
        if (make_tmap_4d(&ta, A, eb, dims, str, box)) return 1;
    }
    {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)Ncols, (uint64_t)batch1, (uint64_t)batch2};
        uint64_t str[3] = {(uint64_t)ldb * eb, (uint64_t)(batch1 > 1 ? b_s1 : ldb * Ncols) * eb,
                           (uint64_t)(batch2 > 1 ? b_s2 : ldb * Ncols * batch1) * eb};
        uint32_t box[4] = {(uint32_t)bk, (uint32_t)(bn / p.cluster), 1, 1};
        if (make_tmap_4d(&tb, B, eb, dims, str, box)) return 1;
    }
    return launch_igemm_nt(dtype, ta, tb, p, bn, (cudaStream_t)stream);
}

// ================================================================================================ attention scores
namespace b2 {
// b2_gemm_nt with B given UN-transposed, [b2][b1][k][n] (row stride ldb): C = alpha * A . B.  The NT kernel consumes it MN-major
// (igemm.h: b_mn), so the attention backward products dV = P^T dO and dK = dS^T Q (autograd of custom_layers.py:144-150) need no
// transposed copies of dO / Q, and the Linear data gradients (dX = dY W, custom_layers.py:116,119) read the forward weight [out][in]
// in place.  bf16 only; Ncols a multiple of 64; always batched addressing (batch1 = batch2 = 1 is fine); residual: unbatched only.
extern "C" int b2_gemm_nt_bmn(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb,
                              long long b_s1, long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2,
                              int M, int Ncols, int K, int batch1, int batch2, float alpha, const void* residual, long long ldr,
                              int dtype, void* stream) {
    if (dtype != 0) return set_error("b2_gemm_nt_bmn: bf16 only");
    if (residual && (batch1 != 1 || batch2 != 1)) return set_error("b2_gemm_nt_bmn: residual only for unbatched GEMM");
    const int eb = 2, bk = 64;
    if (Ncols % 64) return set_error("b2_gemm_nt_bmn: Ncols must be a multiple of 64");
    if ((lda * eb) % 16 || (ldb * eb) % 16 || (b_s1 * eb) % 16 || (b_s2 * eb) % 16) return set_error("b2_gemm_nt_bmn: strides must be 16-byte multiples");
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.W = M; p.H = batch1; p.N = batch2;
    p.wb = 128; p.hb = 1; p.nb = 1;
    p.tiles_w = (M + 127) / 128; p.tiles_h = batch1; p.tiles_n = batch2;
    p.groups = 1; p.taps = 1; p.kb_per_tap = (K + bk - 1) / bk;   // K tail is zero-filled by TMA (both operands)
    p.Cout = Ncols;
    p.out = C; p.oN = c_s2; p.oH = c_s1; p.oW = ldc; p.oC = 1;
    p.vec_ok = ((ldc % 8 == 0) && (c_s1 % 8 == 0) && (c_s2 % 8 == 0) && ((uintptr_t)C % 16 == 0) &&
                (!residual || ((ldr % 8 == 0) && ((uintptr_t)residual % 16 == 0)))) ? 1 : 0;
    p.residual = residual; p.rN = 0; p.rH = 0; p.rW = ldr;
    p.alpha = alpha;
    p.b_mode = 1;
    p.b_mn = 1;
    int bn, splits;
    pick_tiling(Ncols, (long long)p.tiles_w * batch1 * batch2, 1, p.kb_per_tap, false, &bn, &splits);
    p.n_tiles = (Ncols + bn - 1) / bn;
    p.splits = 1; p.ws = workspace(0).ws; p.ws_counters = workspace(0).counters;
    p.cluster = 1;
    CUtensorMap ta, tb;
    {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)batch1, (uint64_t)batch2};
        uint64_t str[3] = {(uint64_t)lda * eb, (uint64_t)(batch1 > 1 ? a_s1 : lda * M) * eb,
                           (uint64_t)(batch2 > 1 ? a_s2 : lda * M * batch1) * eb};
        uint32_t box[4] = {(uint32_t)bk, 128, 1, 1};
        if (make_tmap_4d(&ta, A, eb, dims, str, box)) return 1;
    }
    {
        uint64_t dims[5] = {64, (uint64_t)K, (uint64_t)Ncols / 64, (uint64_t)batch1, (uint64_t)batch2};
        uint64_t str[4] = {(uint64_t)ldb * eb, 128, (uint64_t)(batch1 > 1 ? b_s1 : ldb * K) * eb,
                           (uint64_t)(batch2 > 1 ? b_s2 : ldb * K * batch1) * eb};
        uint32_t box[5] = {64, 64, (uint32_t)(bn / 64), 1, 1};
        if (make_tmap_nd(&tb, B, eb, 5, dims, str, box)) return 1;
    }
    return launch_igemm_nt(dtype, ta, tb, p, bn, (cudaStream_t)stream);
}

int launch_softmax_fixup(void* pt, long long ldp, const float* stats, long long rows, int P, int n_tiles, int tile_cols, int dtype,
                         cudaStream_t st);
}

static int attn_bn(int P) { return P <= 64 ? 64 : (P <= 128 ? 128 : 256); }

// Shared set-up of the two score kernels: rows = keys (A operand), columns = queries (B operand), both [P][d] slices
// of the packed qkv / gradient tensors, batched over (head, image).
static int attn_params(IgemmParams* p, CUtensorMap* ta, CUtensorMap* tb, const void* A, long long lda, long long a_sh,
                       long long a_sn, const void* B, long long ldb, long long b_sh, long long b_sn, void* out, long long ldo,
                       long long o_sh, long long o_sn, int P, int d, int heads, int N, int dtype, int bn) {
    const int eb = dtype == 0 ? 2 : 4;
    const int bk = 128 / eb;
    if ((lda * eb) % 16 || (ldb * eb) % 16) return set_error("attention: row strides must be 16-byte multiples");
    memset(p, 0, sizeof(*p));
    p->W = P; p->H = heads; p->N = N;
    p->wb = 128; p->hb = 1; p->nb = 1;
    p->tiles_w = (P + 127) / 128; p->tiles_h = heads; p->tiles_n = N;
    p->groups = 1; p->taps = 1; p->kb_per_tap = (d + bk - 1) / bk;
    p->Cout = P;
    p->out = out; p->oN = o_sn; p->oH = o_sh; p->oW = ldo; p->oC = 1;
    const int ov = dtype == 1 ? 4 : 8;
    p->vec_ok = ((ldo % ov == 0) && (o_sh % ov == 0) && (o_sn % ov == 0) && ((uintptr_t)out % 16 == 0)) ? 1 : 0;
    p->b_mode = 1;
    p->splits = 1;
    p->cluster = 1;
    p->n_tiles = (P + bn - 1) / bn;
    {
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)P, (uint64_t)heads, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)lda * eb, (uint64_t)(heads > 1 ? a_sh : lda * P) * eb, (uint64_t)(N > 1 ? a_sn : lda * P * heads) * eb};
        uint32_t box[4] = {(uint32_t)bk, 128, 1, 1};
        if (make_tmap_4d(ta, A, eb, dims, str, box)) return 1;
    }
    {
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)P, (uint64_t)heads, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)ldb * eb, (uint64_t)(heads > 1 ? b_sh : ldb * P) * eb, (uint64_t)(N > 1 ? b_sn : ldb * P * heads) * eb};
        uint32_t box[4] = {(uint32_t)bk, (uint32_t)bn, 1, 1};
        if (make_tmap_4d(tb, B, eb, dims, str, box)) return 1;
    }
    return 0;
}

// P^T[n][h][j][i] = softmax over the QUERY index i of scale * q_i . k_j  (custom_layers.py:144-147), computed
// as S^T = K Q^T so that the softmax axis runs along the TMEM columns of one thread: the fp32 score matrix never
// reaches memory.  k, q: [P][d] slices (row stride ld, head stride sh, image stride sn, elements).  pt: row stride ldp
// (multiple of 8 elements).  work: 2 * N*heads*P * ceil(P/256) floats, only touched when P > 256.
extern "C" int b2_attn_scores_softmax(const void* k, const void* q, long long ld, long long sh, long long sn, void* pt,
                                      long long ldp, int P, int d, int heads, int N, float scale, float* work, int dtype,
                                      void* stream) {
    IgemmParams p; CUtensorMap ta, tb;
    const int bn = attn_bn(P);
    if (attn_params(&p, &ta, &tb, k, ld, sh, sn, q, ld, sh, sn, pt, ldp, (long long)P * ldp, (long long)heads * P * ldp, P, d, heads,
                    N, dtype, bn)) return 1;
    p.alpha = scale; p.act = 4;
    if (p.n_tiles > 1) {
        if (!work) return set_error("b2_attn_scores_softmax: P > 256 needs the stats workspace");
        p.gn_stats = work;
    }
    if (launch_igemm_nt(dtype, ta, tb, p, bn, (cudaStream_t)stream)) return 1;
    if (p.n_tiles > 1)
        return launch_softmax_fixup(pt, ldp, work, (long long)N * heads * P, P, p.n_tiles, bn, dtype, (cudaStream_t)stream);
    return 0;
}

// dS^T[n][h][j][i] = scale * P^T[j][i] * (dP^T[j][i] - dot[n][h][j]),  dP^T = V dO^T  (backward of the softmax above;
// dot[j] = sum_i P^T dP^T = sum_c V[j][c] dV[j][c], see b2_rowdot).  v: [P][d] slices of qkv, d_o: [P][d] slices with
// row stride ldo_ / head stride d / image stride P*ldo_.
extern "C" int b2_attn_scores_bwd(const void* v, long long ld, long long sh, long long sn, const void* d_o, long long ld_do,
                                  long long do_sh, long long do_sn, const void* pt, const float* dot, void* dst, long long ldp,
                                  int P, int d, int heads, int N, float scale, int dtype, void* stream) {
    IgemmParams p; CUtensorMap ta, tb;
    const int bn = attn_bn(P);
    if (attn_params(&p, &ta, &tb, v, ld, sh, sn, d_o, ld_do, do_sh, do_sn, dst, ldp, (long long)P * ldp, (long long)heads * P * ldp,
                    P, d, heads, N, dtype, bn)) return 1;
    p.alpha = scale; p.act = 5;
    p.residual = pt; p.rW = ldp; p.rH = (long long)P * ldp; p.rN = (long long)heads * P * ldp;
    p.rowvec = dot; p.vW = heads; p.vH = 1; p.vN = (long long)heads * P;      // dot is [N*P][heads] (b2_rowdot's layout)
    return launch_igemm_nt(dtype, ta, tb, p, bn, (cudaStream_t)stream);
}

// ================================================================================================ TN GEMM launchers
namespace b2 {
int launch_gemm_tn(int dtype, const CUtensorMap& a, const CUtensorMap& b, const GemmTnParams& p, int block_n, cudaStream_t st);
}

// K box of <= 64 pixel rows whose row count is a multiple of the MMA K (16 bf16 / 8 tf32); boxes may overhang the
// tensor (TMA zero-fills), which is what makes tiny images and short sequences legal.
static int pick_kbox(int W, int H, int N, int batch_mode, int kmult, int* wb, int* hb, int* nb) {
    if (batch_mode) {
        int w = W < 64 ? W : 64;
        w = ((w + kmult - 1) / kmult) * kmult;
        *wb = w; *hb = 1; *nb = 1;
        return 0;
    }
    for (int w = (W < 64 ? W : 64); w >= 1; --w) {
        if (W % w) continue;
        for (int h = (H < 64 / w ? H : 64 / w); h >= 1; --h) {
            if (H % h) continue;
            int n = 64 / (w * h);
            if (n > N) n = N;
            if (h < H) n = 1;                       // images may only be batched once whole images fit
            while (n >= 1 && (w * h * n) % kmult) --n;
            if (n >= 1) { *wb = w; *hb = h; *nb = n; return 0; }
            // round the image count UP instead (overhang is zero-filled)
            if (h == H) {
                n = 1;
                while ((w * h * n) % kmult) ++n;
                if (w * h * n <= 64) { *wb = w; *hb = h; *nb = n; return 0; }
            }
        }
    }
    return set_error("gemm_tn: cannot tile a %dx%d image into K boxes of a multiple of %d rows", W, H, kmult);
}

static int tn_block_n(int ncols, int dtype) {
    if (dtype == 0) return ncols >= 256 ? 256 : (ncols >= 128 ? 128 : 64);
    return ncols >= 128 ? 128 : (ncols >= 64 ? 64 : 32);
}

// Split-K factor of the TN kernel: work items = tiles x splits run in waves of (#SMs) persistent CTAs; a split costs
// an fp32 atomic pass over the tile instead of plain stores.  Pick the factor with the least (waves x K boxes per item),
// charging each extra split a few K boxes for the atomics, and prefer no split when the tiles alone fill the machine.
static void tn_pick_splits(GemmTnParams* p, int batches, int bn) {
    const int sms = device_sm_count();
    const long long base = (long long)p->m_tiles * p->n_tiles * p->taps * batches;
    const int k_boxes = p->kt_w * p->kt_h * p->kt_n;
    int best_s = 1;
    if (p->out_mode == 0) {
        double best = 1e30;
        const int cap = k_boxes / 4 > 1 ? k_boxes / 4 : 1;
        for (int s = 1; s <= cap && s <= 64; ++s) {
            const long long items = base * s;
            const long long waves = (items + sms - 1) / sms;
            const double cost = (double)waves * ((k_boxes + s - 1) / s + 6.0) + (s > 1 ? 8.0 + 2.0 * s : 0.0);
            if (cost < best) { best = cost; best_s = s; }
            if (items >= 4LL * sms) break;
        }
    }
    p->splits = best_s;
    // Split-K reduction of the weight gradients.  Default: fp32 vector atomics straight into the gradient buffer.  Ordered
    // (bitwise repeatable: per-split partial tiles, the last arriver sums them in split order) in deterministic mode or with
    // SDM_B200_TN_SPLITK=ordered: MEASURED 10 % slower over the weight gradients of a 128x128 batch-32 step (24.5 vs 22.3 ms;
    // up to 1.6x on the small deep layers, where the hand-off latency is not hidden), hence opt-in.
    p->ws = nullptr; p->ws_counters = nullptr;
    static int ordered_env = -1;
    if (ordered_env < 0) { const char* e = getenv("SDM_B200_TN_SPLITK"); ordered_env = (e && !strcmp(e, "ordered")) ? 1 : 0; }
    if (p->out_mode == 0 && best_s > 1 && (ordered_env || deterministic_mode())) {
        const Workspace& w = workspace(1);
        if (w.ws && base <= 1024) {
            int s = best_s;
            while (s > 1 && base * s * 128LL * bn > w.floats) --s;       // fit the partial tiles, trading splits for capacity
            p->splits = s;
            if (s > 1) { p->ws = w.ws; p->ws_counters = w.counters; }
        }
    }
}

// Plans group g (mode 2: output parity, else 0) of a weight gradient: shapes, tiling, split-K and both operand maps.
static int wgrad_setup(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* dz, int Cout, long long lddz,
                       float* grad_packed, int dtype, int g, GemmTnParams* out, CUtensorMap* ta, CUtensorMap* tb, int* bn_out) {
    const int eb = dtype == 0 ? 2 : 4;
    const int kmult = dtype == 0 ? 16 : 8;
    if (mode < 0 || mode > 2) return set_error("b2_conv2d_wgrad: bad mode %d", mode);
    if ((ldx * eb) % 16 || (lddz * eb) % 16) return set_error("b2_conv2d_wgrad: channel strides must be 16-byte multiples");
    GemmTnParams p;
    memset(&p, 0, sizeof(p));
    p.W = W; p.H = H; p.N = N;
    if (pick_kbox(W, H, N, 0, kmult, &p.wb, &p.hb, &p.nb)) return 1;
    p.kt_w = (W + p.wb - 1) / p.wb; p.kt_h = (H + p.hb - 1) / p.hb; p.kt_n = (N + p.nb - 1) / p.nb;
    p.batch_mode = 0;
    p.M = Cout; p.Ncols = Cin;
    p.m_tiles = (Cout + 127) / 128;
    const int bn = tn_block_n(Cin, dtype);
    p.n_tiles = (Cin + bn - 1) / bn;
    p.out_mode = 0; p.alpha = 1.0f;
    p.tap_stride = Cin;
    const int slab = 128 / eb;
    int b_images = N;
    if (mode == 0 || mode == 1) {
        p.taps = 9; p.ldc = 9LL * Cin;
        for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
            const int t = kh * 3 + kw;
            if (mode == 0) { p.tap_dh[t] = kh - 1; p.tap_dw[t] = kw - 1; p.tap_dn[t] = 0; }
            else {
                const int pr = (kh != 1), pc = (kw != 1);
                p.tap_dh[t] = (kh == 0) ? -1 : 0; p.tap_dw[t] = (kw == 0) ? -1 : 0; p.tap_dn[t] = (pr * 2 + pc) * N;
            }
        }
        if (mode == 1) b_images = 4 * N;
    } else {
        p.taps = 4; p.ldc = 4LL * Cin;
    }
    tn_pick_splits(&p, 1, bn);
    {   // CTA pairs sharing the activation slabs: measured neutral on B200 (weight-gradient layers at batch 32: 5.4 ms either way),
        // so off unless SDM_B200_TN_CLUSTER=2; needs an even number of M tiles and enough items
        static int tn_cl = -1;
        if (tn_cl < 0) { const char* e = getenv("SDM_B200_TN_CLUSTER"); tn_cl = e ? atoi(e) : 1; }
        const long long items = (long long)p.splits * p.m_tiles * p.n_tiles * p.taps;
        p.cluster = (tn_cl == 2 && dtype == 0 && p.m_tiles % 2 == 0 && bn >= 128 && items >= device_sm_count()) ? 2 : 1;
    }
    // slab maps (one TMA instruction per operand box instead of one per 128-byte slab) need whole slabs of channels
    p.box5 = (option("tn_box5", 1) && Cin % slab == 0 && Cout % slab == 0) ? 1 : 0;
    const int a_slabs = 128 / slab, b_slabs = bn / slab;
    const uint32_t spb_b = (uint32_t)tn_spb_b(b_slabs, a_slabs, p.cluster);
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)b_images};
        uint64_t str[3] = {(uint64_t)ldx * eb, (uint64_t)W * ldx * eb, (uint64_t)H * W * ldx * eb};
        uint32_t box[4] = {(uint32_t)slab, (uint32_t)p.wb, (uint32_t)p.hb, (uint32_t)p.nb};
        if (p.box5 ? make_tmap_5d_slabs(tb, x, eb, dims, str, box, spb_b, dtype == 1)
                   : make_tmap_4d(tb, x, eb, dims, str, box, dtype == 1)) return 1;
    }
    {
        const char* dz_base = (const char*)dz;
        uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)lddz * eb, (uint64_t)W * lddz * eb, (uint64_t)H * W * lddz * eb};
        p.out = grad_packed;
        if (mode == 2) {       // dz is [N][2H][2W][Cout]; parity (a, b) is the strided view dz[:, a::2, b::2, :]
            const int a = g >> 1, b = g & 1;
            dz_base += ((long long)a * (2 * W) + b) * lddz * eb;
            str[0] = (uint64_t)2 * lddz * eb; str[1] = (uint64_t)2 * (2 * W) * lddz * eb; str[2] = (uint64_t)4 * H * W * lddz * eb;
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                const int t = i * 2 + j;
                p.tap_dh[t] = (i == 0) ? 0 : (a == 0 ? -1 : 1);
                p.tap_dw[t] = (j == 0) ? 0 : (b == 0 ? -1 : 1);
                p.tap_dn[t] = 0;
            }
            p.out = grad_packed + (long long)g * Cout * 4 * Cin;
        }
        uint32_t box[4] = {(uint32_t)slab, (uint32_t)p.wb, (uint32_t)p.hb, (uint32_t)p.nb};
        if (p.box5 ? make_tmap_5d_slabs(ta, dz_base, eb, dims, str, box, (uint32_t)a_slabs, dtype == 1)
                   : make_tmap_4d(ta, dz_base, eb, dims, str, box, dtype == 1)) return 1;
    }
    *out = p;
    *bn_out = bn;
    return 0;
}

// Weight gradient of the three convolution flavours, accumulated (fp32 atomics) into a zero-initialised buffer in
// KERNEL layout: mode 0/1 -> [Cout][9][Cin], mode 2 -> [4 parities][Cout][4][Cin]  (b2_unpack_weight_grad maps
// it back to the parameter's layout).  x: the forward input (mode 1: its parity planes), dz: gradient w.r.t. the
// conv's pre-activation output.  (H, W): forward A-operand extents, as in b2_conv2d_nhwc.
extern "C" int b2_conv2d_wgrad(int mode, const void* x, int N, int H, int W, int Cin, long long ldx, const void* dz,
                               int Cout, long long lddz, float* grad_packed, int dtype, void* stream) {
    const int n_groups = mode == 2 ? 4 : 1;
    for (int g = 0; g < n_groups; ++g) {
        GemmTnParams p;
        CUtensorMap ta, tb;
        int bn;
        if (wgrad_setup(mode, x, N, H, W, Cin, ldx, dz, Cout, lddz, grad_packed, dtype, g, &p, &ta, &tb, &bn)) return 1;
        if (launch_gemm_tn(dtype, ta, tb, p, bn, (cudaStream_t)stream)) return 1;
    }
    return 0;
}

// Several weight gradients in as few launches as possible (igemm.h: TnJobTable).  desc: n_jobs rows of 11 values
// {mode, x, N, H, W, Cin, ldx, dz, Cout, lddz, grad_packed} (pointers as integers), same meaning as b2_conv2d_wgrad.  Layers the
// grouped kernel cannot take (transposed conv, TF32, channel counts that are not whole 64-channel slabs, ordered split-K, CTA
// pairs) are launched one by one.  The result equals n_jobs calls of b2_conv2d_wgrad (up to the order of fp32 atomic adds).
namespace b2 { int launch_gemm_tn_grouped(const TnJobTable& tab, int block_n, cudaStream_t st); }
extern "C" int b2_conv2d_wgrad_batch(int n_jobs, const long long* desc, int dtype, void* stream) {
    if (n_jobs < 0 || (n_jobs > 0 && !desc)) return set_error("b2_conv2d_wgrad_batch: bad job list");
    static thread_local TnJobTable tabs[3];          // one table per N-tile width (256 / 128 / 64), filled and flushed in turn
    const int widths[3] = {256, 128, 64};
    for (int t = 0; t < 3; ++t) { tabs[t].n_jobs = 0; tabs[t].total_items = 0; }
    auto flush = [&](int t) -> int {
        if (tabs[t].n_jobs == 0) return 0;
        const int rc = launch_gemm_tn_grouped(tabs[t], widths[t], (cudaStream_t)stream);
        tabs[t].n_jobs = 0; tabs[t].total_items = 0;
        return rc;
    };
    for (int i = 0; i < n_jobs; ++i) {
        const long long* d = desc + 11LL * i;
        const int mode = (int)d[0], N = (int)d[2], H = (int)d[3], W = (int)d[4], Cin = (int)d[5], Cout = (int)d[8];
        const void* x = reinterpret_cast<const void*>(d[1]);
        const void* dz = reinterpret_cast<const void*>(d[7]);
        float* grad = reinterpret_cast<float*>(d[10]);
        const long long ldx = d[6], lddz = d[9];
        const int n_groups = mode == 2 ? 4 : 1;
        for (int g = 0; g < n_groups; ++g) {
            GemmTnParams p;
            CUtensorMap ta, tb;
            int bn;
            if (wgrad_setup(mode, x, N, H, W, Cin, ldx, dz, Cout, lddz, grad, dtype, g, &p, &ta, &tb, &bn)) return 1;
            const bool groupable = dtype == 0 && mode != 2 && p.box5 && p.cluster == 1 && p.ws == nullptr && p.taps <= 9 &&
                                   Cin % 64 == 0 && ((uintptr_t)p.out % 16) == 0;
            if (!groupable) {
                if (launch_gemm_tn(dtype, ta, tb, p, bn, (cudaStream_t)stream)) return 1;
                continue;
            }
            const int t = bn == 256 ? 0 : (bn == 128 ? 1 : 2);
            if (tabs[t].n_jobs == kTnMaxJobs && flush(t)) return 1;
            TnJob& J = tabs[t].jobs[tabs[t].n_jobs++];
            J.tmA = ta; J.tmB = tb;
            J.out = reinterpret_cast<float*>(p.out); J.ldc = p.ldc; J.tap_stride = p.tap_stride;
            J.wb = p.wb; J.hb = p.hb; J.nb = p.nb; J.kt_w = p.kt_w; J.kt_h = p.kt_h; J.kt_n = p.kt_n;
            J.m_tiles = p.m_tiles; J.n_tiles = p.n_tiles; J.taps = p.taps; J.splits = p.splits; J.M = p.M; J.Ncols = p.Ncols;
            J.alpha = p.alpha;
            for (int k = 0; k < 9; ++k) { J.tap_dw[k] = p.tap_dw[k]; J.tap_dh[k] = p.tap_dh[k]; J.tap_dn[k] = p.tap_dn[k]; }
            J.item_begin = tabs[t].total_items;
            tabs[t].total_items += p.splits * p.m_tiles * p.n_tiles * p.taps;
        }
    }
    for (int t = 0; t < 3; ++t) if (flush(t)) return 1;
    return 0;
}

// C[b2][b1][m][n] (+)= alpha * sum_k A[b2][b1][k][m] * B[b2][b1][k][n]; A [K][M] (row stride lda), B [K][Ncols].
// out_mode 0: fp32 atomic accumulate into C (Linear weight gradients: dW = dY^T X, k = row);
// out_mode 1: store in `dtype` (attention dV = P^T dO, dK = dS^T Q; batched over head / image).
extern "C" int b2_gemm_tn(const void* A, long long lda, long long a_s1, long long a_s2, const void* B, long long ldb,
                          long long b_s1, long long b_s2, void* C, long long ldc, long long c_s1, long long c_s2, int M,
                          int Ncols, int K, int batch1, int batch2, float alpha, int out_mode, int dtype, void* stream) {
    const int eb = dtype == 0 ? 2 : 4;
    const int kmult = dtype == 0 ? 16 : 8;
    if ((lda * eb) % 16 || (ldb * eb) % 16) return set_error("b2_gemm_tn: lda/ldb rows must be 16-byte multiples");
    GemmTnParams p;
    memset(&p, 0, sizeof(p));
    p.W = K; p.H = batch1; p.N = batch2;
    if (pick_kbox(K, 1, 1, 1, kmult, &p.wb, &p.hb, &p.nb)) return 1;
    p.kt_w = (K + p.wb - 1) / p.wb; p.kt_h = 1; p.kt_n = 1;
    p.batch_mode = 1;
    p.M = M; p.Ncols = Ncols;
    p.m_tiles = (M + 127) / 128;
    const int bn = tn_block_n(Ncols, dtype);
    p.n_tiles = (Ncols + bn - 1) / bn;
    p.taps = 1;
    p.out = C; p.ldc = ldc; p.tap_stride = 0; p.c_s1 = c_s1; p.c_s2 = c_s2;
    p.out_mode = out_mode; p.alpha = alpha;
    tn_pick_splits(&p, batch1 * batch2, bn);
    const int slab = 128 / eb;
    p.box5 = (option("tn_box5", 1) && M % slab == 0 && Ncols % slab == 0) ? 1 : 0;
    CUtensorMap ta, tb;
    {
        uint64_t dims[4] = {(uint64_t)M, (uint64_t)K, (uint64_t)batch1, (uint64_t)batch2};
        uint64_t str[3] = {(uint64_t)lda * eb, (uint64_t)(batch1 > 1 ? a_s1 : lda * K) * eb,
                           (uint64_t)(batch2 > 1 ? a_s2 : lda * K * batch1) * eb};
        uint32_t box[4] = {(uint32_t)slab, (uint32_t)p.wb, 1, 1};
        if (p.box5 ? make_tmap_5d_slabs(&ta, A, eb, dims, str, box, (uint32_t)(128 / slab), dtype == 1)
                   : make_tmap_4d(&ta, A, eb, dims, str, box, dtype == 1)) return 1;
    }
    {
        uint64_t dims[4] = {(uint64_t)Ncols, (uint64_t)K, (uint64_t)batch1, (uint64_t)batch2};
        uint64_t str[3] = {(uint64_t)ldb * eb, (uint64_t)(batch1 > 1 ? b_s1 : ldb * K) * eb,
                           (uint64_t)(batch2 > 1 ? b_s2 : ldb * K * batch1) * eb};
        uint32_t box[4] = {(uint32_t)slab, (uint32_t)p.wb, 1, 1};
        if (p.box5 ? make_tmap_5d_slabs(&tb, B, eb, dims, str, box, (uint32_t)tn_spb_b(bn / slab, 128 / slab, 1), dtype == 1)
                   : make_tmap_4d(&tb, B, eb, dims, str, box, dtype == 1)) return 1;
    }
    return launch_gemm_tn(dtype, ta, tb, p, bn, (cudaStream_t)stream);
}
