// Parameter blocks shared by the tcgen05 implicit-GEMM kernels and their host launchers.
#pragma once
#include <stdint.h>
#include <cuda.h>

namespace b2 {

constexpr int kMaxTaps = 16;   // 9 for 3x3, 4 groups x 4 taps for the 4x4/s2 transposed conv
constexpr int kMaxGroups = 4;

// C[pixel, cout] = sum_{tap, cin} A[pixel + tap offset, cin] * Bw[group][cout][tap*Cin + cin]
// A is reached through a 4-D TMA map (cin | W | H | N); out-of-image taps are zero-filled by TMA.
struct IgemmParams {
    int W, H, N;                 // extents of A-map dims 1..3
    int wb, hb, nb;              // A box over those dims (wb*hb*nb <= 128 rows of the MMA tile)
    int tiles_w, tiles_h, tiles_n;
    int n_tiles;                 // ceil(Cout / BLOCK_N)
    int groups;                  // weight slabs (transposed-conv output parities), else 1
    int taps;                    // taps per group
    int kb_per_tap;              // 128-byte K blocks per tap (= Cin*sizeof(T)/128)
    int b_mode;                  // 0: B batch coords = (group, 0); 1: = (h, n) of the A tile (batched GEMM)
    int tap_dw[kMaxTaps], tap_dh[kMaxTaps], tap_dn[kMaxTaps];   // index: group*taps + tap
    int Cout;                    // valid output columns
    // output addressing (elements): out + n*oN + h*oH + w*oW + goff[group] + column
    void* out;
    long long oN, oH, oW, oC;    // oC: column (channel) stride, 1 for NHWC, H*W for an NCHW result
    int vec_ok;                  // 16-byte vector stores/loads allowed (oC == 1 and everything 16-byte aligned)
    long long goff[kMaxGroups];
    int out_fp32;                // bf16 kernels only: write fp32 instead of bf16
    const void* residual;        // optional, same dtype as out; rN/rH/rW addressing, no group offset
    long long rN, rH, rW;
    const float* bias;           // optional [Cout]
    float alpha;                 // acc *= alpha before bias
    int act;                     // 0 none, 1 swish, 2 tanh, 3 none but GN stats of swish(value),
                                 // 4 softmax over the valid columns of each row (attention: rows = keys, columns = queries;
                                 //   gn_stats then receives per-(row, n-tile) (max, sum) when a row spans several tiles),
                                 // 5 out = alpha * residual[row][col] * (acc - rowvec[row])  (softmax backward)
    // split-K (small-M layers whose few output tiles cannot fill 148 SMs): a work item is (tile, split); every split
    // stores its fp32 partial tile into `ws` ([tile][split][128][BLOCK_N]); the split that arrives last (per-tile counter,
    // self-resetting) sums the slices in split order -- deterministic -- and runs the normal epilogue.
    int cluster;                 // 1, or CL = 2 / 4: CTAs of a cluster share the B tile through TMA multicast (needs splits == 1)
    int splits;
    float* ws;
    int* ws_counters;
    const float* rowvec;         // act 5: one value per output row, index n*vN + h*vH + w*vW
    long long vN, vH, vW;
    float* gn_stats;             // optional [N][Cout/cpg][2] (sum, sum of squares), accumulated atomically
    int cpg;                     // channels per GroupNorm group
    // Halo mode (3x3 stride-1 convs with few input channels: Cin = 128 / 256, whose K loop is too short to amortise the operand
    // traffic).  Instead of re-fetching the 128-pixel A box once per filter tap (9x), ONE box per 64-channel block is fetched
    // that covers the tile plus its halo, and every tap reads it through a descriptor whose start address is shifted by whole
    // 128-byte rows (the SWIZZLE_128B pattern is a function of the absolute shared-memory address, so a row shift is legal:
    // tools/exp_halo_desc.cu).  For the shift to be uniform the M tile is a run of 128 consecutive positions of a FLATTENED
    // image with one shared zero column per row (pitch P = W + 1: position f = h*P + w, w == W is the pad, produced by TMA's
    // out-of-bounds fill and skipped by the epilogue), halo = 1: tiles_w = ceil(H*P / 128), wb = 128;
    // or, for W % 128 == 0, one 128-pixel run of one image row, halo = 2 (box 130 x 3 rows, no pad positions).
    // Column sums for the CONSUMER's GroupNorm x AdaGN backward (b2_conv2d_nhwc_colsum): the data-gradient GEMM that produces
    // `dout` also emits, per (image, channel), cs_s1 += sum_p dout and cs_s2 += sum_p dout * swish(z) from its epilogue, z being
    // the consumer's pre-activation (same pixel indexing as the output).  The backward's reduce pass (a full read of dout and z)
    // disappears; the apply pass turns the raw sums into its a1 / a2 terms.
    const void* cs_z;
    long long cs_zN, cs_zH, cs_zW;
    float* cs_s1;
    float* cs_s2;
    // Swapped-operand mode for Cout <= 128 (BLOCK_N = 256 instance only): the MMA's 128 rows are OUTPUT CHANNELS (A = a 128-row
    // weights tile) and its 256 columns are PIXELS (B = the activation box, wb*hb*nb <= 256), i.e. D^T = W X^T.  A 128-channel
    // layer then runs the 128x256 MMA shape of the wide layers (measured 86-91 % tensor pipe) instead of 128x128 (40-48 %), and
    // the epilogue thread owns one channel: bias is a scalar, GroupNorm sums are in-thread over pixels.  n_tiles = ceil(Cout / 128).
    int swap_ab;
    // b_mn (b2_conv2d_nhwc mode 5: data gradient of the 3x3 stride-1 conv): B is the FORWARD weight tensor [Cout_f][tap][Cin_f]
    // (channels-last storage = the optimiser's bf16 copy) consumed MN-major: a K block is 64 rows (Cout_f) of the mirrored tap,
    // the N tile BLOCK_N/64 slabs of 64 contiguous Cin_f -- one 4-D TMA box (ci | co | ci slab | tap).  No transposed weight copy
    // exists and no per-step transpose kernel runs (104-164 launches of a train step before).
    int b_mn;
    // L2 prefetch of the weight matrix at kernel start (cp.async.bulk.prefetch.L2 over [pf_ptr, pf_ptr + pf_bytes), spread over the
    // CTAs): the K loop touches the matrix in 128-byte pieces that are rows apart (a K block of every output channel), a DRAM-
    // unfriendly pattern when the weights are cold -- and they always are: 1.2 GB of weights pass between two uses of a layer.
    // MEASURED: the just-in-time transposed copies of the data-gradient weights acted as exactly such a prefetch (dropping them
    // without one cost 3.5 ms per 128x128 batch-32 step although the kernels are equally fast in isolation).
    const void* pf_ptr;
    long long pf_bytes;
    // Second output (b2_conv2d_nhwc_dual, act 0): out2 = Swish(value stored to out), same dtype, its own per-pixel stride / group
    // offsets.  Training forward of the un-normalised convs keeps the pre-activation z for backward AND feeds Swish(z) on; the
    // separate Swish pass (read z, write a: 4 bytes per element and one launch per conv) disappears.
    void* out2;
    long long o2N, o2H, o2W;
    long long goff2[kMaxGroups];
    int vec2_ok;
    int halo;                    // 0 off, 1 flattened, 2 row-aligned
    int halo_P;                  // flattened pitch W + 1 (halo 1)
    int halo_msub;               // 128-row sub-tiles per work item (2: both share every weights stage; BLOCK_N = 128 only)
    int halo_BW, halo_R;         // A box: BW columns x R rows of pixels (x 64 channels); shared-memory row = (h - h_lo)*BW + (w - w_lo)
    int a_buf_bytes, a_slots;    // A buffers (one box each), 1024-byte multiples
    int b_stages;                // depth of the B (weights) ring, <= the kernel's STAGES
    int halo_bar_off;            // byte offset of the barrier block behind the A buffers and the B ring
};

// C[m, n] (+)= alpha * sum_k A[k, m] * B[k + tap shift, n]  -- both operands MN-major ("TN" GEMM).
//   weight gradients: k = pixel, m = Cout, n = (tap, Cin)            (reference: autograd's convolution_backward)
//   attention:        k = key / query index, batched over (head, image)
// A and B are reached through 4-D TMA maps (channel | W | H | N); a K block is a box of up to 64 pixel rows.
struct GemmTnParams {
    int W, H, N;                 // pixel extents of the maps' dims 1..3
    int wb, hb, nb;              // K box (wb*hb*nb <= 64 rows)
    int kt_w, kt_h, kt_n;        // K boxes per dim; in batch mode kt_h = kt_n = 1 and (H, N) index the batch
    int batch_mode;              // 0: K spans (w, h, n); 1: K spans w only, tiles are batched over (h, n)
    int m_tiles, n_tiles;        // output tiles: M / 128, ceil(Ncols_per_tap / BLOCK_N)
    int taps;                    // taps (n-tile index = tap * n_tiles_per_tap + j)
    int tap_dw[kMaxTaps], tap_dh[kMaxTaps], tap_dn[kMaxTaps];
    int splits;                  // split-K factor (work item = tile x split)
    int cluster;                 // 1, or 2: CTA pairs on consecutive M tiles share the B (activation) slabs through TMA multicast
    int M, Ncols;                // valid rows / valid columns per tap
    void* out;                   // C; row stride ldc, tap stride tap_stride (elements), batch strides c_s1 (h) / c_s2 (n)
    long long ldc, tap_stride, c_s1, c_s2;
    int out_mode;                // 0: fp32 gradient written into a zero-initialised buffer, 1: store in the operand dtype
    float alpha;
    // out_mode 0 with splits > 1: ws != NULL -> ORDERED split-K (bitwise repeatable): every split stores its fp32 partial tile to
    // ws ([tile][split][128][BLOCK_N]); the split that arrives last (per-tile counter, self-resetting) sums the slices in split
    // order and stores the tile.  ws == NULL -> fp32 vector atomics straight into the output (order-dependent rounding).
    float* ws;
    int* ws_counters;
    // box5: the operand maps are 5-D slab maps (host_util.h: make_tmap_5d_slabs): ONE TMA instruction fetches all slabs of the
    // A operand of a stage and one (or two, BLOCK_N = 256) all slabs of B, instead of one instruction per 128-byte-wide slab.
    // Slabs are then box_rows * 128 bytes apart in shared memory (the descriptor's leading-byte offset follows).
    int box5;
};

// ---- grouped weight gradients (gemm_tn.cu: gemm_tn_grouped_kernel) ------------------------------------------------------------
// At small per-GPU batches the deep U-Net levels turn every weight gradient into a ~16 us launch whose K loop is 3 us: the step is
// bound by the NUMBER of dependent launches.  Weight gradients feed nothing but the optimiser, so the host may defer them and hand
// a whole module's worth to ONE persistent launch: a work item is (job, tile, split), the job table (operand maps + shapes) rides
// in the kernel's parameter space (CUDA 12.1+: up to 32 KB), so nothing has to be uploaded and a captured graph carries it.
constexpr int kTnMaxJobs = 48;
struct alignas(64) TnJob {
    CUtensorMap tmA, tmB;        // 5-D slab maps of dz (A) and x (B)
    float* out;                  // flat gradient slice of the layer, kernel layout [Cout][taps][Cin]
    long long ldc, tap_stride;
    int wb, hb, nb;              // K box (pixels)
    int kt_w, kt_h, kt_n;        // K boxes per dim
    int m_tiles, n_tiles, taps, splits;
    int M, Ncols;
    int item_begin;              // first work item of this job within the launch
    float alpha;
    int tap_dw[9], tap_dh[9], tap_dn[9];
};
struct TnJobTable {
    TnJob jobs[kTnMaxJobs];
    int n_jobs, total_items;
};
static_assert(sizeof(TnJobTable) <= 32000, "the job table must fit the kernel parameter space (32 764 bytes)");

// Slab-map mode of the TN kernel: a B box carries as many slabs as the A box (16 KB) unless a CTA pair has to split the B tile.
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int tn_spb_b(int b_slabs, int a_slabs, int cl) {
    int s = b_slabs / cl;
    if (s > a_slabs) s = a_slabs;
    return s < 1 ? 1 : s;
}

}  // namespace b2
