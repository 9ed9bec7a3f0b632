// tcgen05 implicit-GEMM ("NT": both operands K-major) for sm_100a.
//
// One persistent, warp-specialised kernel serves every dense contraction of the U-Net forward and
// data-gradient paths (reference call sites: models/custom_layers.py:224 Conv2d 3x3, :196 Conv2d 3x3/s2,
// :174 ConvTranspose2d 4x4/s2, :116/:119 attention Linear layers, :144 q.k^T):
//   * A (activations, NHWC) arrives through a 4-D TMA map; a filter tap is just a shifted box, and the
//     zero padding of the convolution is TMA's out-of-bounds fill -- no im2col buffer ever exists.
//   * B (weights, [Cout][tap][Cin]) arrives through a second TMA map.
//   * warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer, warps 2..5 = epilogue
//     (TMEM -> registers -> bias / Swish / GroupNorm partial sums / residual -> global).
//   * accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "ptx.cuh"
#include "igemm.h"
#include "host_util.h"

namespace b2 {

template <int BLOCK_N, int STAGES>
struct IgemmSmem {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = BAR_OFF + (2 * STAGES + 4) * 8 + 16;        // [2 accumulator stages][BLOCK_N] fp32
    // swapped-operand epilogue (BLOCK_N == 256 only): per epilogue warp a [32 pixels][32 channels] bf16 transposition tile, rows
    // padded to 80 bytes (conflict-free 16-byte reads)
    static constexpr int TR_ROW = 40;                                            // halfwords per pixel row
    static constexpr int TR_OFF = (BIAS_OFF + 2 * BLOCK_N * 4 + 15) & ~15;
    static constexpr int TR_BYTES = BLOCK_N == 256 ? 8 * 32 * TR_ROW * 2 : 0;
    static constexpr int TOTAL = TR_OFF + TR_BYTES + 1024;                       // +1024: manual alignment slack
};

struct TileCoord { int nt, w0, h0, n0, g; };

__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
    TileCoord c;
    c.nt = tile % p.n_tiles;  tile /= p.n_tiles;
    c.w0 = (tile % p.tiles_w) * p.wb;  tile /= p.tiles_w;
    c.h0 = (tile % p.tiles_h) * p.hb;  tile /= p.tiles_h;
    c.n0 = (tile % p.tiles_n) * p.nb;  tile /= p.tiles_n;
    c.g = tile;
    return c;
}

// Cluster mode pads the M-tile count to a multiple of CL; a padded (dummy) tile decodes to an image index past the batch.
template <int CL>
__device__ __forceinline__ TileCoord decode_tile_cl(const IgemmParams& p, int tile, int m_groups) {
    if constexpr (CL == 1) {
        return decode_tile(p, tile);
    } else {
        TileCoord c;
        c.nt = tile % p.n_tiles;  tile /= p.n_tiles;
        const int m_pad = m_groups * CL;
        int m = tile % m_pad;
        c.g = tile / m_pad;
        const int m_total = p.tiles_n * p.tiles_h * p.tiles_w;
        if (m >= m_total) { c.w0 = 0; c.h0 = 0; c.n0 = p.tiles_n * p.nb; return c; }     // dummy: every row is out of range
        c.w0 = (m % p.tiles_w) * p.wb;  m /= p.tiles_w;
        c.h0 = (m % p.tiles_h) * p.hb;  m /= p.tiles_h;
        c.n0 = m * p.nb;
        return c;
    }
}

constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int kThreads = 64 + kEpiWarps * 32 + 32;      // + a second TMA producer warp (the last one)

// Per-(image, group) sum / sum-of-squares of one 32-column chunk.  Each lane owns one pixel; the NV = 2*32/GS partial
// values are reduced over the 32 lanes with a reduce-scatter butterfly (NV-1+log2(32/NV) shuffles instead of 5*NV),
// after which NV distinct lanes issue one atomicAdd each.
template <int GS>
__device__ __forceinline__ void gn_partial(const float (&v)[32], bool valid, bool uniform, int lane,
                                           float* stats_row /* stats + n*G*2 */, int group0) {
    constexpr int NG = 32 / GS;
    constexpr int NV = 2 * NG;
    float a[NV];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < GS; ++i) { const float x = v[g * GS + i]; s1 += x; s2 = fmaf(x, x, s2); }
        a[2 * g] = valid ? s1 : 0.f;
        a[2 * g + 1] = valid ? s2 : 0.f;
    }
    if (uniform) {
        int cnt = NV;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            if (cnt > 1) {
                const bool hi = (lane & o) != 0;
                const int half = cnt / 2;
#pragma unroll
                for (int i = 0; i < NV / 2; ++i) {
                    if (i < half) {
                        const float keep = hi ? a[i + half] : a[i];
                        const float send = hi ? a[i] : a[i + half];
                        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
                cnt = half;
            } else {
                a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
            }
        }
        constexpr int LANES_PER_VAL = 32 / NV;
        if ((lane & (LANES_PER_VAL - 1)) == 0) atomicAdd(stats_row + group0 * 2 + lane / LANES_PER_VAL, a[0]);
    } else if (valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) atomicAdd(stats_row + group0 * 2 + i, a[i]);
    }
}

// Column sums over the 32 lanes (pixels) of 32 per-lane values (channels): reduce-scatter butterfly, 31 shuffles; afterwards
// a[0] of lane L is the sum of channel L.
__device__ __forceinline__ void warp_colsum32(float (&a)[32], int lane) {
#pragma unroll
    for (int o = 16, cnt = 32; o > 0; o >>= 1, cnt >>= 1) {
        const bool hi = (lane & o) != 0;
        const int half = cnt / 2;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < half) {
                const float keep = hi ? a[i + half] : a[i];
                const float send = hi ? a[i] : a[i + half];
                a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
    }
}

// CL > 1: thread-block cluster of CL CTAs that work on CL consecutive M tiles of the SAME N tile.  The B (weight) tile
// of every K step is identical for them, so each CTA fetches 1/CL of it and TMA-multicasts that piece to all: the L2 -> SM
// operand traffic per CTA drops from A + B to A + B/CL (the main loop is bound by exactly that traffic, not by the MMA).
// HALO: see IgemmParams::halo -- the A operand of a tile is ONE box per 64-channel block (tile + halo) that all 9 taps read
// through row-shifted descriptors; B (weights) streams through a ring of p.b_stages stages, one stage per (channel block, tap).
constexpr int kMaxASlots = 4;

// B2_NT_MAXREG: cap the registers per thread (all 11 warps get the epilogue warps' allocation: 168 x 352 = 59 136 of the SM's 65 536
// registers, so no other kernel's CTA -- the bucket-wise Adam on its own stream, a side-stream weight gradient's helper kernels --
// can become resident next to a GEMM CTA).
#ifdef B2_NT_MAXREG
#define B2_NT_BOUNDS __maxnreg__(B2_NT_MAXREG)
#else
#define B2_NT_BOUNDS __launch_bounds__(kThreads, 1)
#endif
// B_MN: IgemmParams::b_mn as a COMPILE-TIME switch.  As a run-time one it put the B descriptor, its K step and the instruction
// descriptor into per-thread registers, and the issue loop paid ten R2UR moves per K step (tensor pipe of the wide convs 82 -> 74 %
// under ncu); as a template parameter the descriptors are immediates on the uniform datapath again.
template <typename T, int BLOCK_N, int STAGES, int CL, bool HALO = false, bool B_MN = false>
__global__ void B2_NT_BOUNDS
igemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ IgemmParams p) {
    using S = IgemmSmem<BLOCK_N, STAGES>;
    constexpr bool kTF32 = (sizeof(T) == 4);
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                 : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(2 * BLOCK_N <= 512, "two accumulator stages must fit TMEM");
    // HALO, BLOCK_N = 128: an item is TWO 128-row sub-tiles (p.halo_msub = 2) that share every weights stage, i.e. an accumulator
    // stage is 2 x 128 columns and both stages fill the 512 TMEM columns
    constexpr uint32_t TMEM_ALLOC = HALO ? 512u : TMEM_COLS;
    const int msub = HALO ? p.halo_msub : 1;
    const int acc_cols = msub * BLOCK_N;                 // TMEM columns of one accumulator stage

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* bar_base = HALO ? smem + p.halo_bar_off : smem + S::BAR_OFF;
    uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);
    volatile int* flag_s = reinterpret_cast<volatile int*>(tmem_holder + 1);      // split-K: arrival order of this CTA's split
    uint64_t* afull = tempty + 3;                                                   // HALO: A-buffer barriers
    uint64_t* aempty = afull + kMaxASlots;
    float* bias_s = HALO ? reinterpret_cast<float*>(bar_base + 256) : reinterpret_cast<float*>(smem + S::BIAS_OFF);
    uint8_t* b_ring = smem + (HALO ? p.a_slots * p.a_buf_bytes : 0);                // HALO: weights ring behind the A buffers

    pdl_launch_dependents();               // the next kernel may be scheduled (and run its prologue) while this one works
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        // a stage may be refilled (by multicasts from every CTA of the cluster) once ALL CL consumers released it
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
        if constexpr (HALO) for (int i = 0; i < kMaxASlots; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_ALLOC>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();          // peers' barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                            // everything above overlapped the previous kernel's tail; global memory from here on

    if (warp == 2 && p.pf_bytes > 0) {
        // weights -> L2, 16 KB per instruction, consecutive chunks on different CTAs (igemm.h: pf_ptr); the epilogue warps idle now
        const long long chunks = (p.pf_bytes + 16383) >> 14;
        for (long long c = (long long)lane * gridDim.x + blockIdx.x; c < chunks; c += 32LL * gridDim.x) {
            const long long off = c << 14;
            const long long left = p.pf_bytes - off;
            l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(p.pf_ptr) + off, static_cast<uint32_t>(left < 16384 ? left : 16384));
        }
    }
    const int splits = p.splits;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    // Work items.  CL == 1: (tile, split), CTA i takes items i, i + grid, ...  CL > 1 (no split-K): an item is a group of CL
    // consecutive M tiles x one N tile; cluster c takes groups c, c + #clusters, ... and CTA `crank` the crank-th M tile of the
    // group (tiles past the end are dummies: TMA zero-fills, the epilogue sees no valid row).
    const int m_total = p.tiles_n * p.tiles_h * p.tiles_w;
    const int m_groups = (m_total + CL - 1) / CL;
    const int total_tiles = CL > 1 ? p.groups * m_groups * p.n_tiles
                                   : p.groups * m_total * p.n_tiles * splits;
    const int item0 = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int item_step = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;
    auto tile_of = [&](int item) -> int {              // linear tile index in decode_tile()'s order (N tile fastest)
        if constexpr (CL > 1) {
            const int nt = item % p.n_tiles;
            const int rest = item / p.n_tiles;
            const int mg = rest % m_groups, g = rest / m_groups;
            return (g * (m_groups * CL) + mg * CL + crank) * p.n_tiles + nt;       // may exceed the real tile count: dummy
        } else {
            return item / splits;
        }
    };
    const int k_iters = p.taps * p.kb_per_tap;
    constexpr int BK_ELEMS = 128 / sizeof(T);
    const uint32_t a_bytes = static_cast<uint32_t>(p.wb * p.hb * p.nb) * 128u;

    if (warp == 0 || warp == kThreads / 32 - 1) {
        // ------------------------------------------------------------ TMA producers
        // Two of them: warp 0 fetches the A (activation) box and posts the stage's byte count, the last warp fetches B.
        // One thread issuing both boxes of every stage could not keep the ring full on the short-K layers (ncu source
        // sampling: the producer sat on UTMALDG while the MMA warp waited for data; same finding as in gemm_tn.cu).
        const bool load_a = warp == 0;
        if constexpr (HALO) {
            if (load_a) {
                // ---- A: one box per (tile, 64-channel block): the tile's 128 positions plus their halo
                // (producer loops are warp-uniform too: all lanes wait, one elected lane issues -- coordinates stay uniform)
                int slot = 0; uint32_t aph = 0;
                const uint32_t box_bytes = static_cast<uint32_t>(p.halo_BW * p.halo_R) * 128u;
                for (int item = item0; item < total_tiles; item += item_step) {
                    const TileCoord tc = decode_tile_cl<CL>(p, tile_of(item), m_groups);
                    int w_lo, h_lo;
                    if (p.halo == 1) {                    // flattened: rows floor((f0 - P - 1) / P) .. of the image, all P columns
                        const int t = tc.w0 - p.halo_P - 1;
                        h_lo = t >= 0 ? t / p.halo_P : -((-t + p.halo_P - 1) / p.halo_P);
                        w_lo = 0;
                    } else {                              // row-aligned: rows h-1 .. h+1, columns w0-1 .. w0+128
                        h_lo = tc.h0 - 1; w_lo = tc.w0 - 1;
                    }
                    for (int cb = 0; cb < p.kb_per_tap; ++cb) {
                        mbar_wait(&aempty[slot], aph ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&afull[slot], box_bytes);
                            tma_load_4d(smem + slot * p.a_buf_bytes, &tmA, &afull[slot], cb * 64, w_lo, h_lo, tc.n0);
                        }
                        __syncwarp();
                        if (++slot == p.a_slots) { slot = 0; aph ^= 1; }
                    }
                }
            } else {
                // ---- B: one stage per (64-channel block, tap), K coordinate tap*Cin + cb*64 of the packed weights
                int s = 0; uint32_t ph = 0;
                for (int item = item0; item < total_tiles; item += item_step) {
                    const TileCoord tc = decode_tile_cl<CL>(p, tile_of(item), m_groups);
                    for (int cb = 0; cb < p.kb_per_tap; ++cb)
                        for (int t = 0; t < 9; ++t) {
                            const int it = t * p.kb_per_tap + cb;
                            mbar_wait(&empty[s], ph ^ 1);
                            uint8_t* b_dst = b_ring + s * S::B_BYTES;
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&full[s], S::B_BYTES);
                                if constexpr (B_MN) {     // forward-layout weights, MN-major (IgemmParams::b_mn): (ci | co | ci slab | tap)
                                    if constexpr (CL > 1) {
                                        constexpr int PIECE = BLOCK_N / CL;
                                        if constexpr (PIECE >= 64)
                                            tma_load_4d_mc(b_dst + crank * PIECE * 128, &tmB, &full[s], 0, cb * 64,
                                                           tc.nt * (BLOCK_N / 64) + crank * (PIECE / 64), 8 - t, kMask);
                                    } else {
                                        tma_load_4d(b_dst, &tmB, &full[s], 0, cb * 64, tc.nt * (BLOCK_N / 64), 8 - t);
                                    }
                                } else if constexpr (CL > 1) {
                                    constexpr int PIECE = BLOCK_N / CL;
                                    tma_load_4d_mc(b_dst + crank * PIECE * 128, &tmB, &full[s], it * 64, tc.nt * BLOCK_N + crank * PIECE, 0, 0, kMask);
                                } else {
                                    tma_load_4d(b_dst, &tmB, &full[s], it * 64, tc.nt * BLOCK_N, 0, 0);
                                }
                            }
                            __syncwarp();
                            if (++s == p.b_stages) { s = 0; ph ^= 1; }
                        }
                }
            }
        } else
        {
            int s = 0; uint32_t ph = 0;
            for (int item = item0; item < total_tiles; item += item_step) {
                const int split = CL > 1 ? 0 : item % splits;
                const TileCoord tc = decode_tile_cl<CL>(p, tile_of(item), m_groups);
                const int bb2 = p.b_mode ? tc.h0 : tc.g;
                const int bb3 = p.b_mode ? tc.n0 : 0;
                const int it0 = (int)(((long long)k_iters * split) / splits), it1 = (int)(((long long)k_iters * (split + 1)) / splits);
                int t = it0 / p.kb_per_tap, kb = it0 - t * p.kb_per_tap;       // (tap, K block) advance without a division per K step
                // the tap's box origin changes once per kb_per_tap K steps: looked up then, not per step (the indexed constant loads
                // were a third of the A producer's samples in the ncu source page of the 1024-channel conv)
                int aw = 0, ah = 0, an = 0;
                auto tap_origin = [&]() {
                    const int ti = tc.g * p.taps + t;
                    aw = tc.w0 + p.tap_dw[ti]; ah = tc.h0 + p.tap_dh[ti]; an = tc.n0 + p.tap_dn[ti];
                };
                if (it0 < it1) tap_origin();
                for (int it = it0; it < it1; ++it) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* a_dst = smem + s * S::STAGE_BYTES;
                    uint8_t* b_dst = a_dst + S::A_BYTES;
                    if (elect_one()) {
                        if (p.swap_ab) {
                            // swapped operands: the 16 KB A region takes the 128-row weights tile, the B region the pixel box
                            if (load_a) {
                                mbar_arrive_expect_tx(&full[s], a_bytes + S::A_BYTES);
                                tma_load_4d(a_dst, &tmB, &full[s], it * BK_ELEMS, tc.nt * 128, bb2, bb3);
                            } else {
                                tma_load_4d(b_dst, &tmA, &full[s], kb * BK_ELEMS, aw, ah, an);
                            }
                        } else if (load_a) {
                            mbar_arrive_expect_tx(&full[s], a_bytes + S::B_BYTES);
                            tma_load_4d(a_dst, &tmA, &full[s], kb * BK_ELEMS, aw, ah, an);
                        } else if constexpr (B_MN) {
                            // data gradient straight from the FORWARD weights [Cout][tap][Cin] (IgemmParams::b_mn): the K block is 64
                            // output channels (rows) of the mirrored tap, the N tile BLOCK_N/64 slabs of 64 input channels
                            const int wt = p.taps - 1 - t;
                            if constexpr (CL > 1) {
                                constexpr int PIECE = BLOCK_N / CL;
                                if constexpr (PIECE >= 64)
                                    tma_load_4d_mc(b_dst + crank * PIECE * 128, &tmB, &full[s], 0, kb * 64,
                                                   tc.nt * (BLOCK_N / 64) + crank * (PIECE / 64), wt, kMask);
                            } else if (p.b_mode) {
                                // batched GEMM with B given as [K][Ncols] (b2_gemm_nt_bmn): (n | k | n slab | batch 1 | batch 2)
                                tma_load_5d(b_dst, &tmB, &full[s], 0, kb * 64, tc.nt * (BLOCK_N / 64), bb2, bb3);
                            } else {
                                tma_load_4d(b_dst, &tmB, &full[s], 0, kb * 64, tc.nt * (BLOCK_N / 64), wt);
                            }
                        } else if constexpr (CL > 1) {
                            constexpr int PIECE = BLOCK_N / CL;         // rows of the B tile this CTA fetches for the whole cluster
                            tma_load_4d_mc(b_dst + crank * PIECE * 128, &tmB, &full[s], it * BK_ELEMS, tc.nt * BLOCK_N + crank * PIECE,
                                           bb2, bb3, kMask);
                        } else {
                            tma_load_4d(b_dst, &tmB, &full[s], it * BK_ELEMS, tc.nt * BLOCK_N, bb2, bb3);
                        }
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                    if (++kb == p.kb_per_tap) { kb = 0; ++t; if (it + 1 < it1) tap_origin(); }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if constexpr (HALO) {
            {
                // b_mn: the weights stage is BLOCK_N/64 MN-major slabs of [64 k rows][128 B] (8 KB apart); a 16-row K step is 2 KB
                constexpr uint32_t idesc = umma_idesc(1u, 128, BLOCK_N, 0, B_MN ? 1u : 0u);
                const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);          // descriptor without its address field
                const uint64_t bdesc0 = B_MN ? umma_desc_sw128(0, 64 * 128, 1024) : desc0;
                constexpr uint32_t bk_step = B_MN ? 128u : 2u;
                const uint32_t smem0 = smem_u32(smem) & 0x3FFFFu;      // CTA-local offset (see ptx.cuh: umma_desc_sw128)
                const uint32_t ring0 = smem_u32(b_ring) & 0x3FFFFu;
                int s = 0; uint32_t ph = 0;
                int slot = 0; uint32_t aph = 0;
                int acc = 0; uint32_t acc_ph = 0;
                for (int item = item0; item < total_tiles; item += item_step) {
                    const TileCoord tc = decode_tile_cl<CL>(p, tile_of(item), m_groups);
                    int row0;                              // shared-memory row of the tile's position 0 under the centre tap
                    if (p.halo == 1) {
                        const int t = tc.w0 - p.halo_P - 1;
                        const int h_lo = t >= 0 ? t / p.halo_P : -((-t + p.halo_P - 1) / p.halo_P);
                        row0 = tc.w0 - h_lo * p.halo_P;
                    } else {
                        row0 = p.halo_BW + 1;
                    }
                    mbar_wait(&tempty[acc], acc_ph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * acc_cols;
                    const uint32_t sub_rows = p.halo == 1 ? 128u : static_cast<uint32_t>(p.halo_BW);   // rows between the two sub-tiles
                    for (int cb = 0; cb < p.kb_per_tap; ++cb) {
                        mbar_wait(&afull[slot], aph);
                        tc_fence_after();
                        const uint32_t a_base = smem0 + slot * p.a_buf_bytes;
                        for (int t = 0; t < 9; ++t) {
                            const int dh = t / 3 - 1, dw = t % 3 - 1;
                            mbar_wait(&full[s], ph);
                            tc_fence_after();
                            const uint64_t ad0 = desc0 + ((a_base + static_cast<uint32_t>(row0 + dh * p.halo_BW + dw) * 128u) >> 4);
                            const uint64_t bd0 = bdesc0 + ((ring0 + s * S::B_BYTES) >> 4);
                            if (elect_one()) {
                                for (int sub = 0; sub < msub; ++sub) {            // the sub-tiles share this weights stage (interleaving their MMAs k-outer measured 12 % slower)
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_ss<false>(d_tmem + sub * BLOCK_N, ad0 + sub * (sub_rows * 8u) + 2 * k, bd0 + bk_step * k, idesc,
                                                       (cb > 0 || t > 0 || k != 0) ? 1u : 0u);
                                }
                                if constexpr (CL > 1) umma_commit_mc(&empty[s], kMask); else umma_commit(&empty[s]);
                            }
                            __syncwarp();
                            if (++s == p.b_stages) { s = 0; ph ^= 1; }
                        }
                        if (elect_one()) umma_commit(&aempty[slot]);          // every MMA that read this box has completed
                        __syncwarp();
                        if (++slot == p.a_slots) { slot = 0; aph ^= 1; }
                    }
                    if (elect_one()) umma_commit(&tfull[acc]);
                    __syncwarp();
                    if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                }
            }
        } else
        {
            // warp-uniform issue loop (see elect_one): all lanes wait and compute, one elected lane issues
            constexpr uint32_t idesc = umma_idesc(kTF32 ? 2u : 1u, 128, BLOCK_N, 0, B_MN ? 1u : 0u);
            const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);              // descriptor without its address field
            const uint64_t bdesc0 = B_MN ? umma_desc_sw128(0, 64 * 128, 1024) : desc0;       // b_mn: MN-major slabs (see the halo branch)
            constexpr uint32_t bk_step = B_MN ? 128u : 2u;
            const uint32_t smem0 = smem_u32(smem) & 0x3FFFFu;      // CTA-local offset (see ptx.cuh: umma_desc_sw128)
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            for (int item = item0; item < total_tiles; item += item_step) {
                const int split = CL > 1 ? 0 : item % splits;
                const int it0 = (int)(((long long)k_iters * split) / splits), it1 = (int)(((long long)k_iters * (split + 1)) / splits);
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int it = it0; it < it1; ++it) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t ad0 = desc0 + ((smem0 + s * S::STAGE_BYTES) >> 4);
                    const uint64_t bd0 = B_MN ? bdesc0 + ((smem0 + s * S::STAGE_BYTES + S::A_BYTES) >> 4) : ad0 + (S::A_BYTES >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_ss<kTF32>(d_tmem, ad0 + 2 * k, bd0 + bk_step * k, idesc, (it > it0 || k != 0) ? 1u : 0u);
                        if constexpr (CL > 1) umma_commit_mc(&empty[s], kMask); else umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit(&tfull[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        const int q = warp & 3;                    // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;          // which interleaved set of 32-column chunks this warp owns
        const int epi_tid = threadIdx.x - 64;
        const int row = q * 32 + lane;
        const int w_in = row % p.wb;
        const int h_in = (row / p.wb) % p.hb;
        const int n_in = row / (p.wb * p.hb);
        const int G = p.gn_stats ? (p.Cout / p.cpg) : 0;
        const bool f32out = kTF32 || p.out_fp32;
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = item0; item < total_tiles; item += item_step)
        for (int sub = 0; sub < msub; ++sub) {
            const int tile = tile_of(item);
            TileCoord tc = decode_tile_cl<CL>(p, tile, m_groups);
            if (HALO && sub) { if (p.halo == 1) tc.w0 += 128; else tc.h0 += 1; }      // second 128-row sub-tile of the item
            const uint32_t acc_col = static_cast<uint32_t>(acc * acc_cols + sub * BLOCK_N);
            int w = tc.w0 + w_in, h = tc.h0 + h_in;
            const int n = tc.n0 + n_in;
            if (HALO && p.halo == 1) {                  // flattened position f = h*P + w; w == W is the shared pad column
                const int f = w;
                h = f / p.halo_P;
                w = f - h * p.halo_P;
            }
            const bool valid = (n_in < p.nb) && (w < p.W) && (h < p.H) && (n < p.N);
            const long long o_off = n * p.oN + h * p.oH + w * p.oW + p.goff[tc.g];
            const long long r_off = n * p.rN + h * p.rH + w * p.rW;
            const int n_lane0 = __shfl_sync(0xffffffffu, n, 0);
            // HALO: a tile never leaves its image, so the warp-wide reduction is always legal (pad positions contribute zeros)
            const bool uniform = HALO ? (n_lane0 < p.N)
                                      : (__all_sync(0xffffffffu, n == n_lane0 || !valid) && __shfl_sync(0xffffffffu, (int)valid, 0));
            // stage this tile's bias slice once (all epilogue warps), then meet at a named barrier
            float* bs = bias_s + (acc & 1) * BLOCK_N;
            if (epi_tid < BLOCK_N) {
                const int c = tc.nt * BLOCK_N + epi_tid;
                bs[epi_tid] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");

            mbar_wait(&tfull[acc], acc_ph);
            tc_fence_after();
            if (p.swap_ab) {
                // ---- swapped operands: TMEM lane = output channel, columns = the tile's pixels.  Per 32-pixel chunk: bias /
                // activation / GroupNorm sums with lane = channel (sums are in-thread over pixels), then a 32x32 transposition
                // through shared memory so that every lane stores the 64 contiguous bytes of ONE pixel (16-byte vectors).
                if constexpr (BLOCK_N == 256 && !kTF32 && !HALO) {
                    const int c0w = tc.nt * 128 + q * 32;                // first channel of this warp's lane quarter
                    const int c = c0w + lane;
                    const bool warp_ok = c0w < p.Cout;                   // Cout is a multiple of 32 in this mode
                    const float bias_c = (p.bias && warp_ok) ? __ldg(p.bias + c) : 0.f;
                    const int red_lanes = p.cpg < 32 ? p.cpg : 32;       // lanes (channels) of one GroupNorm group inside this warp
                    __nv_bfloat16* tr = reinterpret_cast<__nv_bfloat16*>(smem + S::TR_OFF) + (warp - 2) * (32 * S::TR_ROW);
                    const int box_px = p.wb * p.hb;                      // pixels per image inside the box (a multiple of 32)
#pragma unroll 1
                    for (int ch = half; ch < 8; ch += 2) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col + ch * 32, r);
                        tmem_ld_wait();
                        if (!warp_ok) continue;
                        // this lane's pixel (after the transposition) and the chunk's image
                        const int pl = ch * 32 + lane;
                        const int pn = pl / box_px, prem = pl - pn * box_px;
                        const int w2 = tc.w0 + prem % p.wb, h2 = tc.h0 + prem / p.wb, n2 = tc.n0 + pn;
                        const bool pv = (pn < p.nb) && (w2 < p.W) && (h2 < p.H) && (n2 < p.N);
                        const uint32_t vmask = __ballot_sync(0xffffffffu, pv);
                        if (vmask == 0u) continue;
                        const int n_chunk = __shfl_sync(0xffffffffu, n2, __ffs(vmask) - 1);   // one image per chunk (box_px % 32 == 0)
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float v = fmaf(__uint_as_float(r[j]), p.alpha, bias_c);
                            if (p.act == 1) v = swish_fast(v);
                            else if (p.act == 2) v = tanhf(v);
                            const float sv = p.act == 3 ? swish_fast(v) : v;     // training: statistics of swish(z), z is stored
                            if ((vmask >> j) & 1u) { s1 += sv; s2 = fmaf(sv, sv, s2); }
                            tr[j * S::TR_ROW + lane] = __float2bfloat16(v);
                        }
                        if (p.gn_stats) {
                            for (int o = 1; o < red_lanes; o <<= 1) {
                                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                            }
                            if ((lane & (red_lanes - 1)) == 0) {
                                float* st = p.gn_stats + ((long long)n_chunk * G + c / p.cpg) * 2;
                                atomicAdd(st, s1);
                                atomicAdd(st + 1, s2);
                            }
                        }
                        __syncwarp();
                        uint4 x[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const uint4*>(tr + lane * S::TR_ROW + i * 8);
                        __syncwarp();
                        if (pv) {
                            const long long off = n2 * p.oN + h2 * p.oH + w2 * p.oW + p.goff[tc.g] + c0w;
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
                            if (p.residual) {
                                const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(p.residual) + n2 * p.rN + h2 * p.rH + w2 * p.rW + c0w;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    uint4 y;
                                    if (p.vec_ok) y = __ldg(reinterpret_cast<const uint4*>(rs) + i);
                                    else { __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(&y); for (int e = 0; e < 8; ++e) yy[e] = rs[i * 8 + e]; }
                                    __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&x[i]);
                                    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 fa = __bfloat1622float2(a2[e]), fb = __bfloat1622float2(b2[e]);
                                        a2[e] = __floats2bfloat162_rn(fa.x + fb.x, fa.y + fb.y);
                                    }
                                }
                            }
                            if (p.vec_ok) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(o)[i] = x[i];
                            } else {
                                const __nv_bfloat16* xx = reinterpret_cast<const __nv_bfloat16*>(x);
                                for (int e = 0; e < 32; ++e) o[e] = xx[e];
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                continue;
            }
            bool from_ws = false;
            float* ws_row = nullptr;
            if (splits > 1) {
                // ---- split-K: publish this split's partial sums, then only the last split to arrive runs the epilogue
                ws_row = p.ws + (((long long)tile * splits) * 128 + row) * BLOCK_N;          // slice of split 0; split s is s*128*BLOCK_N further
                float* mine = ws_row + (long long)(item % splits) * 128 * BLOCK_N;
#pragma unroll 1
                for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
                    if (tc.nt * BLOCK_N + ch * 32 >= p.Cout) break;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col + ch * 32, r);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            __stcg(reinterpret_cast<float4*>(mine + ch * 32) + i,
                                   make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                               __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
                    }
                }
                tc_fence_before();
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
                if (epi_tid == 0) {
                    mbar_arrive_n(&tempty[acc], kEpiWarps);           // the accumulator stage is free again
                    *flag_s = atomicAdd(p.ws_counters + tile, 1);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
                const bool last = (*flag_s == splits - 1);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                if (!last) continue;
                __threadfence();
                from_ws = true;
                if (epi_tid == 0) p.ws_counters[tile] = 0;
            }
            if (p.act == 4) {
                // ---- attention scores: each thread owns one key row of S^T (columns = queries), so the reference's
                // softmax over the QUERY axis (custom_layers.py:147) is a purely in-thread reduction over TMEM columns.
                if (half == 0) {
                    const int ncols_tile = min(BLOCK_N, p.Cout - tc.nt * BLOCK_N);
                    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col;
                    // exp(x) = exp2(x * log2 e).  The scaled score is rounded ONCE (__fmul_rn, never contracted into an FMA
                    // with the subtraction): the row maximum must map to exactly exp2(0) even for logits of 1e14, which
                    // the first DDIM steps of the cosine schedule produce on an untrained net (1/sqrt(abar_T) = 2e7).
                    const float sc = p.alpha * 1.4426950408889634f;
                    float mx = -INFINITY;
#pragma unroll 1
                    for (int c0 = 0; c0 < ncols_tile; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(trow + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (c0 + i < ncols_tile) mx = fmaxf(mx, __fmul_rn(__uint_as_float(r[i]), sc));
                    }
                    float z = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < ncols_tile; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(trow + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (c0 + i < ncols_tile) z += exp2f(__fsub_rn(__fmul_rn(__uint_as_float(r[i]), sc), mx));
                    }
                    const float inv = p.n_tiles == 1 ? 1.0f / z : 1.0f;       // several tiles per row: normalised by the fix-up pass
#pragma unroll 1
                    for (int c0 = 0; c0 < ncols_tile; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(trow + c0, r);
                        tmem_ld_wait();
                        if (!valid) continue;
                        const int col0 = tc.nt * BLOCK_N + c0;
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = exp2f(__fsub_rn(__fmul_rn(__uint_as_float(r[i]), sc), mx)) * inv;
                        const bool full = p.vec_ok && (c0 + 32 <= ncols_tile);
                        if (f32out) {
                            float* o = reinterpret_cast<float*>(p.out) + o_off + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (c0 + i < ncols_tile) o[i] = kTF32 ? round_tf32(v[i]) : v[i];
                        } else {
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + o_off + col0;
                            if (full) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    uint4 x;
                                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
                                    reinterpret_cast<uint4*>(o)[i] = x;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) if (c0 + i < ncols_tile) o[i] = __float2bfloat16(v[i]);
                            }
                        }
                    }
                    if (valid && p.n_tiles > 1) {
                        float* st = p.gn_stats + ((((long long)n * p.H + h) * p.W + w) * p.n_tiles + tc.nt) * 2;
                        st[0] = mx; st[1] = z;
                    }
                }
            } else
#pragma unroll 1
            for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
                const int col0 = tc.nt * BLOCK_N + ch * 32;
                if (col0 >= p.Cout) break;
                uint32_t r[32];
                if (from_ws) {          // sum the partial tiles of all splits (plain L2 reads, fixed order: deterministic)
                    float4 a4[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) {
                        for (int sp = 0; sp < splits; ++sp) {
                            const float4* src = reinterpret_cast<const float4*>(ws_row + (long long)sp * 128 * BLOCK_N + ch * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 x = __ldcg(src + i);
                                a4[i].x += x.x; a4[i].y += x.y; a4[i].z += x.z; a4[i].w += x.w;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        r[4 * i] = __float_as_uint(a4[i].x); r[4 * i + 1] = __float_as_uint(a4[i].y);
                        r[4 * i + 2] = __float_as_uint(a4[i].z); r[4 * i + 3] = __float_as_uint(a4[i].w);
                    }
                } else {
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col + ch * 32, r);
                    tmem_ld_wait();
                }
                const int ncols = min(32, p.Cout - col0);
                const bool vec = p.vec_ok && ncols == 32;
                float v[32];
                const float4* b4 = reinterpret_cast<const float4*>(bs + ch * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = b4[i];
                    v[4 * i + 0] = fmaf(__uint_as_float(r[4 * i + 0]), p.alpha, b.x);
                    v[4 * i + 1] = fmaf(__uint_as_float(r[4 * i + 1]), p.alpha, b.y);
                    v[4 * i + 2] = fmaf(__uint_as_float(r[4 * i + 2]), p.alpha, b.z);
                    v[4 * i + 3] = fmaf(__uint_as_float(r[4 * i + 3]), p.alpha, b.w);
                }
                if (p.act == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = swish_t<!kTF32>(v[i]);
                } else if (p.act == 2) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
                }
                if (p.act == 5) {       // softmax backward: dS^T = alpha * P^T .* (dP^T - dot[row]); P^T arrives as `residual`
                    const float rv = valid ? __ldg(p.rowvec + n * p.vN + h * p.vH + w * p.vW) : 0.f;
                    if (valid) {
                        if (kTF32) {
                            const float* ms = reinterpret_cast<const float*>(p.residual) + r_off + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) v[i] = (v[i] - p.alpha * rv) * ms[i];
                        } else {
                            const __nv_bfloat16* ms = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_off + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) v[i] = (v[i] - p.alpha * rv) * __bfloat162float(ms[i]);
                        }
                    }
                }
                if (p.gn_stats && ncols == 32 && p.act != 5) {
                    float* srow = p.gn_stats + static_cast<long long>(uniform ? n_lane0 : (valid ? n : 0)) * G * 2;
                    const int cpg = p.cpg;
                    if (p.act == 3) {       // training: store the pre-activation, normalise Swish(z) later -> stats of Swish(z)
                        float sw[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) sw[i] = swish_t<!kTF32>(v[i]);
                        if (cpg >= 32)      gn_partial<32>(sw, valid, uniform, lane, srow, col0 / cpg);
                        else if (cpg == 16) gn_partial<16>(sw, valid, uniform, lane, srow, col0 / 16);
                        else if (cpg == 8)  gn_partial<8>(sw, valid, uniform, lane, srow, col0 / 8);
                        else                gn_partial<4>(sw, valid, uniform, lane, srow, col0 / 4);
                    } else {
                        if (cpg >= 32)      gn_partial<32>(v, valid, uniform, lane, srow, col0 / cpg);
                        else if (cpg == 16) gn_partial<16>(v, valid, uniform, lane, srow, col0 / 16);
                        else if (cpg == 8)  gn_partial<8>(v, valid, uniform, lane, srow, col0 / 8);
                        else                gn_partial<4>(v, valid, uniform, lane, srow, col0 / 4);
                    }
                }
                if (valid) {
                    if (f32out) {
                        float* o = reinterpret_cast<float*>(p.out) + o_off + col0 * p.oC;
                        if (p.residual && p.act != 5) {
                            const float* rs = reinterpret_cast<const float*>(p.residual) + r_off + col0;
                            if (vec) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    float4 x = __ldg(reinterpret_cast<const float4*>(rs) + i);
                                    v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += rs[i];
                            }
                        }
                        if (kTF32 && !p.out_fp32) {   // feeds the next kind::tf32 MMA: round to nearest, don't truncate
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
                        }
                        if (vec) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) o[i * p.oC] = v[i];
                        }
                    } else {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + o_off + col0 * p.oC;
                        if (p.residual && p.act != 5) {
                            const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_off + col0;
                            if (vec) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    uint4 x = __ldg(reinterpret_cast<const uint4*>(rs) + i);
                                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        float2 f = __bfloat1622float2(h2[j]);
                                        v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __bfloat162float(rs[i]);
                            }
                        }
                        if (vec) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 x;
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
                                for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
                                reinterpret_cast<uint4*>(o)[i] = x;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) o[i * p.oC] = __float2bfloat16(v[i]);
                        }
                    }
                }
                if (p.out2 && valid) {
                    // second output: Swish of the value just stored (rounded to the storage type first, so that the result equals
                    // the separate Swish pass over the stored pre-activation bit for bit)
                    const long long o2_off = n * p.o2N + h * p.o2H + w * p.o2W + p.goff2[tc.g] + col0;
                    const bool vec2 = p.vec2_ok && ncols == 32;
                    if (f32out) {
                        float* o2 = reinterpret_cast<float*>(p.out2) + o2_off;
                        float a[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) { const float sw = swish_t<!kTF32>(v[i]); a[i] = (kTF32 && !p.out_fp32) ? round_tf32(sw) : sw; }
                        if (vec2) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(o2)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) o2[i] = a[i];
                        }
                    } else {
                        __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) + o2_off;
                        float a[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) a[i] = swish_t<true>(__bfloat162float(__float2bfloat16(v[i])));
                        if (vec2) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 x;
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
                                for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(a[8 * i + 2 * j], a[8 * i + 2 * j + 1]);
                                reinterpret_cast<uint4*>(o2)[i] = x;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) o2[i] = __float2bfloat16(a[i]);
                        }
                    }
                }
                if (p.cs_s1 && !f32out && ncols == 32) {
                    // ---- column sums for the consumer's AdaGN backward (IgemmParams::cs_*): v holds this row's 32 final values
                    float a[32];
                    if (valid) {
                        const __nv_bfloat16* zs = reinterpret_cast<const __nv_bfloat16*>(p.cs_z) + n * p.cs_zN + h * p.cs_zH + w * p.cs_zW + col0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 x = __ldg(reinterpret_cast<const uint4*>(zs) + i);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = __bfloat1622float2(h2[j]);
                                a[8 * i + 2 * j] = v[8 * i + 2 * j] * swish_fast(f.x);
                                a[8 * i + 2 * j + 1] = v[8 * i + 2 * j + 1] * swish_fast(f.y);
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { a[i] = 0.f; v[i] = 0.f; }
                    }
                    if (uniform) {
                        warp_colsum32(a, lane);
                        warp_colsum32(v, lane);
                        float* s1 = p.cs_s1 + (long long)n_lane0 * p.Cout + col0 + lane;
                        float* s2 = p.cs_s2 + (long long)n_lane0 * p.Cout + col0 + lane;
                        atomicAdd(s1, v[0]);
                        atomicAdd(s2, a[0]);
                    } else if (valid) {
#pragma unroll 1
                        for (int i = 0; i < 32; ++i) {
                            atomicAdd(p.cs_s1 + (long long)n * p.Cout + col0 + i, v[i]);
                            atomicAdd(p.cs_s2 + (long long)n * p.Cout + col0 + i, a[i]);
                        }
                    }
                }
            }
            if (splits > 1) continue;                 // accumulator already released right after the partials were published
            tc_fence_before();
            __syncwarp();
            if (sub == msub - 1) {                    // the accumulator stage is free once its last sub-tile has been read
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it / signal its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_ALLOC>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ host
template <typename T, int BLOCK_N, int STAGES, int CL, bool HALO = false, bool B_MN = false>
static int launch_cfg(const CUtensorMap& a, const CUtensorMap& b, const IgemmParams& p, int num_sms, cudaStream_t st) {
    using S = IgemmSmem<BLOCK_N, STAGES>;
    auto kern = igemm_nt_kernel<T, BLOCK_N, STAGES, CL, HALO, B_MN>;
    static bool attr_set = false;
    static int max_clusters = 0;
    constexpr int kHaloMax = 227 * 1024;               // halo mode sizes its buffers per layer: opt in to the maximum once
    const int smem_bytes = HALO ? p.halo_bar_off + 256 + 2 * BLOCK_N * 4 + 1024 : S::TOTAL;
    if (HALO && (smem_bytes > kHaloMax || p.b_stages > STAGES || p.a_slots > kMaxASlots || p.a_slots < 1 || p.b_stages < 2))
        return set_error("igemm_nt (halo): buffers do not fit (%d bytes, %d A slots, %d B stages)", smem_bytes, p.a_slots, p.b_stages);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (CL > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = CL; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO ? kHaloMax : S::TOTAL);
        if (e != cudaSuccess) return set_error("igemm_nt: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        if (CL > 1) {
            // clusters must sit inside one GPC: ask how many fit on this device with one 200 KB CTA per SM
            cfg.gridDim = dim3((num_sms / CL) * CL);
            int n = 0;
            e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
            if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = num_sms / CL; }
            max_clusters = n;
        }
        attr_set = true;
    }
    int grid;
    if (CL > 1) {
        const int m_total = p.tiles_n * p.tiles_h * p.tiles_w;
        const long long groups = (long long)p.groups * ((m_total + CL - 1) / CL) * p.n_tiles;
        const long long cl_cap = max_clusters < num_sms / CL ? max_clusters : num_sms / CL;      // honours the sm_limit option
        const long long clusters = groups < cl_cap ? groups : cl_cap;
        grid = (int)clusters * CL;
    } else {
        const int total_tiles = p.groups * p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles * p.splits;
        grid = total_tiles < num_sms ? total_tiles : num_sms;
    }
    cfg.gridDim = dim3(grid);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, b, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return set_error("igemm_nt launch: %s", cudaGetErrorString(e)); }
    return 0;
}

int launch_igemm_nt(int dtype /*0 bf16, 1 fp32(tf32)*/, const CUtensorMap& a, const CUtensorMap& b,
                    const IgemmParams& p, int block_n, cudaStream_t st) {
    const int sms = device_sm_count();
    const int cl = p.cluster > 1 ? p.cluster : 1;
    if (cl > 1 && (p.splits != 1 || p.b_mode || p.act == 4)) return set_error("igemm_nt: cluster multicast needs an unsplit, unbatched GEMM");
    if (p.swap_ab && (dtype != 0 || block_n != 256 || cl != 1 || p.splits != 1 || p.halo || p.out_fp32))
        return set_error("igemm_nt: swapped-operand mode needs the bf16 256-column kernel without cluster / split-K");
    if (p.b_mn && (dtype != 0 || (p.b_mode && cl != 1) || p.swap_ab || block_n / 64 < cl || p.Cout % 64))
        return set_error("igemm_nt: MN-major weights need the bf16 kernel, whole 64-channel slabs and cluster <= block_n / 64");
    if (p.b_mn) {           // MN-major forward weights: separate instantiations (see igemm_nt_kernel: B_MN)
        if (p.halo) {
            if (p.splits != 1 || p.taps != 9) return set_error("igemm_nt (halo): bf16 3x3 stride-1 convolutions only");
            if (cl == 1 && block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 1, true, true>(a, b, p, sms, st);
            if (cl == 2 && block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 2, true, true>(a, b, p, sms, st);
            return set_error("igemm_nt (halo, MN-major weights): unsupported block_n %d / cluster %d", block_n, cl);
        }
        if (cl == 1) {
            if (block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 1, false, true>(a, b, p, sms, st);
            if (block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 1, false, true>(a, b, p, sms, st);
            if (block_n == 64)  return launch_cfg<__nv_bfloat16, 64, 8, 1, false, true>(a, b, p, sms, st);
        } else if (cl == 2) {
            if (block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 2, false, true>(a, b, p, sms, st);
            if (block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 2, false, true>(a, b, p, sms, st);
        }
        return set_error("igemm_nt (MN-major weights): unsupported block_n %d / cluster %d", block_n, cl);
    }
    if (p.halo) {
        if (dtype != 0 || p.splits != 1 || p.b_mode || p.taps != 9) return set_error("igemm_nt (halo): bf16 3x3 stride-1 convolutions only");
        if (cl == 1 && block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 1, true>(a, b, p, sms, st);
        if (cl == 1 && block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 1, true>(a, b, p, sms, st);
        if (cl == 2 && block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 2, true>(a, b, p, sms, st);
        if (cl == 2 && block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 2, true>(a, b, p, sms, st);
        return set_error("igemm_nt (halo): unsupported block_n %d / cluster %d", block_n, cl);
    }
    if (dtype == 0) {
        if (cl == 1) {
            if (block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 1>(a, b, p, sms, st);
            if (block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 1>(a, b, p, sms, st);
            if (block_n == 64)  return launch_cfg<__nv_bfloat16, 64, 8, 1>(a, b, p, sms, st);
        } else if (cl == 2) {
            if (block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 2>(a, b, p, sms, st);
            if (block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 2>(a, b, p, sms, st);
        } else if (cl == 4) {
            if (block_n == 256) return launch_cfg<__nv_bfloat16, 256, 4, 4>(a, b, p, sms, st);
            if (block_n == 128) return launch_cfg<__nv_bfloat16, 128, 6, 4>(a, b, p, sms, st);
        }
    } else if (cl == 1) {
        if (block_n == 256) return launch_cfg<float, 256, 4, 1>(a, b, p, sms, st);
        if (block_n == 128) return launch_cfg<float, 128, 6, 1>(a, b, p, sms, st);
        if (block_n == 64)  return launch_cfg<float, 64, 8, 1>(a, b, p, sms, st);
    }
    return set_error("igemm_nt: unsupported block_n %d / cluster %d for dtype %d", block_n, cl, dtype);
}

}  // namespace b2
