// tcgen05 "TN" GEMM for sm_100a: C[m][n] (+)= alpha * sum_k A[k][m] * B[k][n], both operands MN-major in shared
// memory (the contraction index is the row index of both global tensors).  This is the shape of every weight
// gradient of the U-Net (k = pixel: dW[cout][tap][cin] = sum_p dz[p][cout] * x[p + tap][cin], the reference's
// convolution_backward / addmm backward) and of the attention products contracted over the sequence axis
// (dV = P^T dO, dK = dS^T Q; reference models/custom_layers.py:144-150 under autograd).
//   * operands arrive through 4-D TMA maps (channel | W | H | N); a filter tap is a shifted box with TMA zero fill,
//     so the im2col matrix of the weight gradient never exists;
//   * split-K over pixel boxes with fp32 vector atomics straight into the (flat) gradient buffer;
//   * same warp roles / TMEM double buffering as igemm_nt.cu.
#include "ptx.cuh"
#include "igemm.h"
#include "host_util.h"

namespace b2 {

constexpr int kTnEpiWarps = 8;
#ifndef TN_PRODUCERS
#define TN_PRODUCERS 3
#endif
constexpr int kTnMaxProducers = TN_PRODUCERS;             // TMA producer warps (slab sl is issued by producer sl % TN_PRODUCERS); measured 2 / 3 / 6: 927 / 953 / 906 TFLOP/s
constexpr int kTnThreads = 64 + kTnEpiWarps * 32 + (kTnMaxProducers - 1) * 32;   // warp 0 + warps 10.. : TMA producers, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int kTnBK = 64;     // pixel rows per pipeline stage

template <int BLOCK_N, int STAGES>
struct TnSmem {
    static constexpr int A_BYTES = 2 * kTnBK * 128;                 // two 64-wide M slabs
    static constexpr int B_BYTES = (BLOCK_N / 64) * kTnBK * 128;    // BLOCK_N/64 slabs (bf16) -- see slab() for fp32
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;
};

struct TnWork { int mt, nt_in_tap, tap, split, bh, bn; };

__device__ __forceinline__ TnWork tn_decode(const GemmTnParams& p, int item) {
    TnWork w;
    w.split = item % p.splits;  item /= p.splits;
    w.mt = item % p.m_tiles;    item /= p.m_tiles;
    w.nt_in_tap = item % p.n_tiles;  item /= p.n_tiles;
    w.tap = item % p.taps;      item /= p.taps;
    if (p.batch_mode) { w.bh = item % p.H; w.bn = item / p.H; } else { w.bh = 0; w.bn = 0; }
    return w;
}

// T = bf16: 64 elements per 128-byte slab row; T = float (tf32): 32 elements per slab row.
// CL == 2: the two CTAs of a cluster take M tiles 2q and 2q+1 of the same (split, N tile, tap); the B slabs of a stage are
// identical for them, so each fetches half of the slabs and TMA-multicasts them to both (same scheme as igemm_nt.cu).
// ORDERED: a separate instantiation, so that the default (atomic) kernel carries none of the ordered epilogue's code or
// registers (measured: the mere presence of the branch cost the atomic path 6 %).
template <typename T, int BLOCK_N, int STAGES, int CL, bool ORDERED = false>
__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmTnParams p) {
    constexpr bool kTF32 = (sizeof(T) == 4);
    constexpr int SLAB = 128 / sizeof(T);                 // MN elements per slab
    constexpr int A_SLABS = 128 / SLAB;
    constexpr int B_SLABS = BLOCK_N / SLAB;
    constexpr int SLAB_BYTES = kTnBK * 128;
    constexpr int A_BYTES = A_SLABS * SLAB_BYTES;
    constexpr int B_BYTES = B_SLABS * SLAB_BYTES;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int UMMA_K = 32 / sizeof(T);                // 16 (bf16) / 8 (tf32)
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
    // slab-map mode (p.box5): slabs per B box / number of B boxes (igemm.h: tn_spb_b, shared with the host)
    constexpr int SPB_B = tn_spb_b(B_SLABS, A_SLABS, CL);
    constexpr int NB_B = B_SLABS / SPB_B;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);
    volatile int* flag_s = reinterpret_cast<volatile int*>(tmem_holder + 1);      // ordered split-K: arrival order of this CTA's split

    pdl_launch_dependents();               // the next kernel may be scheduled (and run its prologue) while this one works
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kTnEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                            // everything above overlapped the previous kernel's tail; global memory from here on

    const int batches = p.batch_mode ? p.H * p.N : 1;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    // CL == 2 (never batched): an item is a PAIR of M tiles; cluster c takes pairs c, c + #clusters, ...
    const int total_items = CL > 1 ? p.splits * (p.m_tiles / CL) * p.n_tiles * p.taps
                                   : p.splits * p.m_tiles * p.n_tiles * p.taps * batches;
    const int item0 = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int item_step = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;
    auto decode = [&](int item) -> TnWork {
        if constexpr (CL > 1) {
            TnWork w;
            w.split = item % p.splits;  item /= p.splits;
            const int mp = item % (p.m_tiles / CL);  item /= (p.m_tiles / CL);
            w.mt = mp * CL + crank;
            w.nt_in_tap = item % p.n_tiles;  item /= p.n_tiles;
            w.tap = item;
            w.bh = 0; w.bn = 0;
            return w;
        } else {
            return tn_decode(p, item);
        }
    };
    const int k_boxes = p.kt_w * p.kt_h * p.kt_n;
    const uint32_t box_rows = static_cast<uint32_t>(p.wb * p.hb * p.nb);
    const uint32_t slab_stride = p.box5 ? box_rows * 128u : static_cast<uint32_t>(SLAB_BYTES);     // bytes between MN slabs in smem

    if (warp == 0 || warp >= 2 + kTnEpiWarps) {
        // TMA producer warps (lane 0 of each): a stage is 2 + BLOCK_N/64 slab loads (one 128-byte-wide box each: the
        // SWIZZLE_128B limit), and a single thread could not issue them fast enough to keep the 4-stage ring full (ncu: the
        // producer sat on UTMALDG while the MMA warp starved).  Producer q takes the slabs sl with sl % n_prod == q; producer 0
        // posts the stage's byte count.  MEASURED alternative, rejected: one warp issuing one slab per LANE ran 40 % slower
        // (wgrad 970 -> 600 TFLOP/s at batch 32) -- TMA issue from several lanes of one warp serialises.
        constexpr int n_prod = kTnMaxProducers;       // compile-time: a runtime count put an integer division per slab into the issue loop (-35 %)
        const int prod = warp == 0 ? 0 : warp - (2 + kTnEpiWarps) + 1;
        if (prod < n_prod) {       // warp-uniform loop: all lanes wait, one elected lane issues (ptx.cuh: elect_one)
            int s = 0; uint32_t ph = 0;
            for (int item = item0; item < total_items; item += item_step) {
                const TnWork wk = decode(item);
                const int kb0 = (int)(((long long)k_boxes * wk.split) / p.splits);
                const int kb1 = (int)(((long long)k_boxes * (wk.split + 1)) / p.splits);
                // K-box coordinates advance incrementally: the divisions by kt_w / kt_h that decode kb cost ~8 % of this kernel's
                // issue slots when they sat inside the loop (ncu source page, round 2) -- in the warps whose pace feeds the MMAs
                int iw = kb0 % p.kt_w, ih = (kb0 / p.kt_w) % p.kt_h, in_ = kb0 / (p.kt_w * p.kt_h);
                const int kt_w = p.kt_w, kt_h = p.kt_h, wb = p.wb, hb = p.hb, nb = p.nb;
                const bool batched = p.batch_mode != 0;
                const int dw = p.tap_dw[wk.tap], dh = batched ? 0 : p.tap_dh[wk.tap], dn = batched ? 0 : p.tap_dn[wk.tap];
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int w0 = iw * wb;
                    const int h0 = batched ? wk.bh : ih * hb;
                    const int n0 = batched ? wk.bn : in_ * nb;
                    if (++iw == kt_w) { iw = 0; if (++ih == kt_h) { ih = 0; ++in_; } }
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* a_dst = smem + s * STAGE_BYTES;
                    uint8_t* b_dst = a_dst + A_BYTES;
                    const int bw = w0 + dw, bh = h0 + dh, bn = n0 + dn;
                    if (elect_one()) {
                    if (prod == 0) mbar_arrive_expect_tx(&full[s], (A_SLABS + B_SLABS) * box_rows * 128u);
                    if (p.box5) {
                        // slab maps: the whole A operand of the stage is one box (producer 0), B is NB_B boxes of SPB_B slabs
                        // (producers 1, 2, ...): at most one TMA instruction per producer warp and stage
                        if (prod == 0) tma_load_5d(a_dst, &tmA, &full[s], 0, w0, h0, n0, wk.mt * A_SLABS);
#pragma unroll
                        for (int j = 0; j < NB_B; ++j) {
                            if (((1 + j) % n_prod) != prod) continue;
                            uint8_t* dst = b_dst + j * SPB_B * slab_stride;
                            const int sl0 = wk.nt_in_tap * B_SLABS + j * SPB_B;
                            if constexpr (CL > 1) {
                                if ((j % CL) != crank) continue;
                                tma_load_5d_mc(dst, &tmB, &full[s], 0, bw, bh, bn, sl0, kMask);
                            } else {
                                tma_load_5d(dst, &tmB, &full[s], 0, bw, bh, bn, sl0);
                            }
                        }
                    } else {
#pragma unroll
                    for (int sl = 0; sl < A_SLABS + B_SLABS; ++sl) {
                        if ((sl % n_prod) != prod) continue;
                        if (sl < A_SLABS) {
                            tma_load_4d(a_dst + sl * SLAB_BYTES, &tmA, &full[s], wk.mt * 128 + sl * SLAB, w0, h0, n0);
                        } else {
                            const int bs = sl - A_SLABS;
                            if constexpr (CL > 1) {
                                // each CTA of the pair fetches half of the B slabs and multicasts them to both
                                if ((bs % CL) != crank) continue;
                                tma_load_4d_mc(b_dst + bs * SLAB_BYTES, &tmB, &full[s], wk.nt_in_tap * BLOCK_N + bs * SLAB, bw, bh, bn, kMask);
                            } else {
                                tma_load_4d(b_dst + bs * SLAB_BYTES, &tmB, &full[s], wk.nt_in_tap * BLOCK_N + bs * SLAB, bw, bh, bn);
                            }
                        }
                    }
                    }
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {
            // warp-uniform issue loop (ptx.cuh: elect_one): all lanes wait and compute the descriptors, one elected lane issues
            constexpr uint32_t idesc = umma_idesc(kTF32 ? 2u : 1u, 128, BLOCK_N, 1, 1);
            // bf16: 8-row atoms (1024 B) of 16-byte chunks; tf32: 4-row atoms (512 B) of 32-byte chunks
            const uint64_t desc0 = umma_desc_sw128(0, slab_stride, kTF32 ? 512 : 1024, kTF32 ? 1 : 2);
            const uint32_t smem0 = smem_u32(smem) & 0x3FFFFu;      // CTA-local offset (see ptx.cuh: umma_desc_sw128)
            constexpr uint32_t K_STEP = (UMMA_K * 128) >> 4;                 // descriptor units (16 B) per MMA K step
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            const int k_steps = (int)(box_rows + UMMA_K - 1) / UMMA_K;      // rows beyond the box are never touched
            for (int item = item0; item < total_items; item += item_step) {
                const TnWork wk = decode(item);
                const int kb0 = (int)(((long long)k_boxes * wk.split) / p.splits);
                const int kb1 = (int)(((long long)k_boxes * (wk.split + 1)) / p.splits);
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t ad0 = desc0 + ((smem0 + s * STAGE_BYTES) >> 4);
                    const uint64_t bd0 = ad0 + (A_BYTES >> 4);
                    if (elect_one()) {
                        constexpr int MAX_K = kTnBK / UMMA_K;
#pragma unroll
                        for (int k = 0; k < MAX_K; ++k)
                            if (k < k_steps) umma_ss<kTF32>(d_tmem, ad0 + k * K_STEP, bd0 + k * K_STEP, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
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
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int epi_tid = threadIdx.x - 64;
        constexpr bool ordered = ORDERED;
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = item0; item < total_items; item += item_step) {
            const TnWork wk = decode(item);
            const int m = wk.mt * 128 + q * 32 + lane;
            const bool valid = m < p.M;
            const long long off = (long long)m * p.ldc + (long long)wk.tap * p.tap_stride + wk.bh * p.c_s1 + wk.bn * p.c_s2;
            const int kb0 = (int)(((long long)k_boxes * wk.split) / p.splits);
            const int kb1 = (int)(((long long)k_boxes * (wk.split + 1)) / p.splits);
            mbar_wait(&tfull[acc], acc_ph);
            tc_fence_after();
            if constexpr (ordered) {
                // ---- ordered split-K: publish this split's partial tile, then only the split that arrives last reduces
                // (never batched, so the tile index is the item index without its split digit -- also in cluster mode, where
                // the two CTAs of a pair see the same item but different M tiles: give each its own tile id)
                const int tile = (item / p.splits) * CL + (CL > 1 ? crank : 0);
                float* ws_row = p.ws + (((long long)tile * p.splits) * 128 + (q * 32 + lane)) * BLOCK_N;
                float* mine = ws_row + (long long)wk.split * 128 * BLOCK_N;
#pragma unroll 1
                for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
                    if (wk.nt_in_tap * BLOCK_N + ch * 32 >= p.Ncols) break;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + ch * 32, r);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            __stcg(reinterpret_cast<float4*>(mine + ch * 32) + i,
                                   kb1 > kb0 ? make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                           __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]))
                                             : make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
                tc_fence_before();
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(kTnEpiWarps * 32) : "memory");
                if (epi_tid == 0) {
                    mbar_arrive_n(&tempty[acc], kTnEpiWarps);         // the accumulator stage is free again
                    *flag_s = atomicAdd(p.ws_counters + tile, 1);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTnEpiWarps * 32) : "memory");
                const bool last = (*flag_s == p.splits - 1);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
                if (!last) continue;
                __threadfence();
                if (epi_tid == 0) p.ws_counters[tile] = 0;
#pragma unroll 1
                for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
                    const int col0 = wk.nt_in_tap * BLOCK_N + ch * 32;
                    if (col0 >= p.Ncols) break;
                    if (!valid) continue;
                    float4 a4[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int sp = 0; sp < p.splits; ++sp) {            // fixed order: bitwise repeatable
                        const float4* src = reinterpret_cast<const float4*>(ws_row + (long long)sp * 128 * BLOCK_N + ch * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 x = __ldcg(src + i);
                            a4[i].x += x.x; a4[i].y += x.y; a4[i].z += x.z; a4[i].w += x.w;
                        }
                    }
                    const int ncols = min(32, p.Ncols - col0);
                    float* o = reinterpret_cast<float*>(p.out) + off + col0;
                    if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(o)[i] = make_float4(a4[i].x * p.alpha, a4[i].y * p.alpha, a4[i].z * p.alpha, a4[i].w * p.alpha);
                    } else {
                        const float* f = reinterpret_cast<const float*>(a4);
                        for (int i = 0; i < ncols; ++i) o[i] = f[i] * p.alpha;
                    }
                }
                continue;
            }
            if (kb1 > kb0) {
#pragma unroll 1
                for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
                    const int col0 = wk.nt_in_tap * BLOCK_N + ch * 32;
                    if (col0 >= p.Ncols) break;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + ch * 32, r);
                    tmem_ld_wait();
                    const int ncols = min(32, p.Ncols - col0);
                    if (!valid) continue;
                    if (p.out_mode == 0 && p.splits == 1) {
                        // one work item owns this output tile: the buffer is zero on entry (ABI), so a plain store IS the
                        // accumulation -- no atomics (they throttled the epilogue below the main loop's pace)
                        float* o = reinterpret_cast<float*>(p.out) + off + col0;
                        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                reinterpret_cast<float4*>(o)[i] =
                                    make_float4(__uint_as_float(r[4 * i]) * p.alpha, __uint_as_float(r[4 * i + 1]) * p.alpha,
                                                __uint_as_float(r[4 * i + 2]) * p.alpha, __uint_as_float(r[4 * i + 3]) * p.alpha);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = __uint_as_float(r[i]) * p.alpha;
                        }
                    } else if (p.out_mode == 0) {
                        float* o = reinterpret_cast<float*>(p.out) + off + col0;
                        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                atomicAdd(reinterpret_cast<float4*>(o) + i,
                                          make_float4(__uint_as_float(r[4 * i]) * p.alpha, __uint_as_float(r[4 * i + 1]) * p.alpha,
                                                      __uint_as_float(r[4 * i + 2]) * p.alpha, __uint_as_float(r[4 * i + 3]) * p.alpha));
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) atomicAdd(o + i, __uint_as_float(r[i]) * p.alpha);
                        }
                    } else if (kTF32) {
                        float* o = reinterpret_cast<float*>(p.out) + off + col0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = round_tf32(__uint_as_float(r[i]) * p.alpha);
                    } else {
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off + col0;
                        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 x;
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&x);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    h2[j] = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2 * j]) * p.alpha,
                                                                  __uint_as_float(r[8 * i + 2 * j + 1]) * p.alpha);
                                reinterpret_cast<uint4*>(o)[i] = x;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = __float2bfloat16(__uint_as_float(r[i]) * p.alpha);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ grouped weight gradients
// Same pipeline as gemm_tn_kernel (bf16, no cluster, atomic split-K, slab maps), but a work item carries a JOB index: the
// operand maps and shapes of up to kTnMaxJobs layers live in the kernel's parameter space (igemm.h: TnJobTable).  CTA i takes
// items i, i + grid, ...; the three roles walk the same item sequence and look the job up with a running index.
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_grouped_kernel(const __grid_constant__ TnJobTable tab) {
    constexpr int SLAB = 64, A_SLABS = 2, B_SLABS = BLOCK_N / SLAB;
    constexpr int SLAB_BYTES = kTnBK * 128;
    constexpr int A_BYTES = A_SLABS * SLAB_BYTES, B_BYTES = B_SLABS * SLAB_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int UMMA_K = 16;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
    constexpr int SPB_B = tn_spb_b(B_SLABS, A_SLABS, 1);
    constexpr int NB_B = B_SLABS / SPB_B;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kTnEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();

    const int total_items = tab.total_items;
    const int n_jobs = tab.n_jobs;
    struct Work { int mt, nt, tap, kb0, kb1; };
    auto advance = [&](int item, int& j) { while (j + 1 < n_jobs && item >= tab.jobs[j + 1].item_begin) ++j; };
    auto decode = [&](const TnJob& J, int item) -> Work {
        Work w;
        int local = item - J.item_begin;
        const int split = local % J.splits;  local /= J.splits;
        w.mt = local % J.m_tiles;            local /= J.m_tiles;
        w.nt = local % J.n_tiles;            local /= J.n_tiles;
        w.tap = local;
        const int k_boxes = J.kt_w * J.kt_h * J.kt_n;
        w.kb0 = (int)(((long long)k_boxes * split) / J.splits);
        w.kb1 = (int)(((long long)k_boxes * (split + 1)) / J.splits);
        return w;
    };

    if (warp == 0 || warp >= 2 + kTnEpiWarps) {
        constexpr int n_prod = kTnMaxProducers;
        const int prod = warp == 0 ? 0 : warp - (2 + kTnEpiWarps) + 1;
        if (prod < n_prod) {
            int s = 0; uint32_t ph = 0;
            int j = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                advance(item, j);
                const TnJob& J = tab.jobs[j];
                const Work wk = decode(J, item);
                const uint32_t box_rows = static_cast<uint32_t>(J.wb * J.hb * J.nb);
                const uint32_t slab_stride = box_rows * 128u;
                const int dw = J.tap_dw[wk.tap], dh = J.tap_dh[wk.tap], dn = J.tap_dn[wk.tap];
                int iw = wk.kb0 % J.kt_w, ih = (wk.kb0 / J.kt_w) % J.kt_h, in_ = wk.kb0 / (J.kt_w * J.kt_h);
                const int kt_w = J.kt_w, kt_h = J.kt_h, wb = J.wb, hb = J.hb, nb = J.nb;
                for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
                    const int w0 = iw * wb, h0 = ih * hb, n0 = in_ * nb;
                    if (++iw == kt_w) { iw = 0; if (++ih == kt_h) { ih = 0; ++in_; } }
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* a_dst = smem + s * STAGE_BYTES;
                    uint8_t* b_dst = a_dst + A_BYTES;
                    if (elect_one()) {
                        if (prod == 0) {
                            mbar_arrive_expect_tx(&full[s], (A_SLABS + B_SLABS) * box_rows * 128u);
                            tma_load_5d(a_dst, &J.tmA, &full[s], 0, w0, h0, n0, wk.mt * A_SLABS);
                        }
#pragma unroll
                        for (int b = 0; b < NB_B; ++b) {
                            if (((1 + b) % n_prod) != prod) continue;
                            tma_load_5d(b_dst + b * SPB_B * slab_stride, &J.tmB, &full[s], 0, w0 + dw, h0 + dh, n0 + dn,
                                        wk.nt * B_SLABS + b * SPB_B);
                        }
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(1u, 128, BLOCK_N, 1, 1);
        const uint32_t smem0 = smem_u32(smem) & 0x3FFFFu;
        constexpr uint32_t K_STEP = (UMMA_K * 128) >> 4;
        int s = 0; uint32_t ph = 0;
        int acc = 0; uint32_t acc_ph = 0;
        int j = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            advance(item, j);
            const TnJob& J = tab.jobs[j];
            const Work wk = decode(J, item);
            const uint32_t box_rows = static_cast<uint32_t>(J.wb * J.hb * J.nb);
            const int k_steps = (int)(box_rows + UMMA_K - 1) / UMMA_K;
            const uint64_t desc0 = umma_desc_sw128(0, box_rows * 128u, 1024, 2);
            mbar_wait(&tempty[acc], acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t ad0 = desc0 + ((smem0 + s * STAGE_BYTES) >> 4);
                const uint64_t bd0 = ad0 + (A_BYTES >> 4);
                if (elect_one()) {
                    constexpr int MAX_K = kTnBK / UMMA_K;
#pragma unroll
                    for (int k = 0; k < MAX_K; ++k)
                        if (k < k_steps) umma_ss<false>(d_tmem, ad0 + k * K_STEP, bd0 + k * K_STEP, idesc, (kb > wk.kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_ph = 0;
        int j = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            advance(item, j);
            const TnJob& J = tab.jobs[j];
            const Work wk = decode(J, item);
            const int m = wk.mt * 128 + q * 32 + lane;
            const bool valid = m < J.M;
            const long long off = (long long)m * J.ldc + (long long)wk.tap * J.tap_stride;
            const float alpha = J.alpha;
            const bool atomic = J.splits > 1;
            mbar_wait(&tfull[acc], acc_ph);
            tc_fence_after();
            if (wk.kb1 > wk.kb0) {
#pragma unroll 1
                for (int ch = half; ch < BLOCK_N / 32; ch += 2) {
                    const int col0 = wk.nt * BLOCK_N + ch * 32;
                    if (col0 >= J.Ncols) break;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + ch * 32, r);
                    tmem_ld_wait();
                    if (!valid) continue;
                    float* o = J.out + off + col0;              // Ncols is a multiple of 64 and every slice 16-byte aligned (host check)
                    if (!atomic) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(o)[i] =
                                make_float4(__uint_as_float(r[4 * i]) * alpha, __uint_as_float(r[4 * i + 1]) * alpha,
                                            __uint_as_float(r[4 * i + 2]) * alpha, __uint_as_float(r[4 * i + 3]) * alpha);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            atomicAdd(reinterpret_cast<float4*>(o) + i,
                                      make_float4(__uint_as_float(r[4 * i]) * alpha, __uint_as_float(r[4 * i + 1]) * alpha,
                                                  __uint_as_float(r[4 * i + 2]) * alpha, __uint_as_float(r[4 * i + 3]) * alpha));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int BLOCK_N, int STAGES>
static int launch_tn_grouped_cfg(const TnJobTable& tab, int num_sms, cudaStream_t st) {
    constexpr int stage_bytes = (2 + BLOCK_N / 64) * kTnBK * 128;
    constexpr int total = STAGES * stage_bytes + (2 * STAGES + 4) * 8 + 16 + 1024;
    static_assert(total <= 227 * 1024, "shared memory budget");
    auto kern = gemm_tn_grouped_kernel<BLOCK_N, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, total);
        if (e != cudaSuccess) return set_error("gemm_tn (grouped): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kTnThreads);
    cfg.dynamicSmemBytes = total;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    int na = 0;
    if (pdl_enabled()) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    cfg.gridDim = dim3(tab.total_items < num_sms ? tab.total_items : num_sms);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tab);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return set_error("gemm_tn (grouped) launch: %s", cudaGetErrorString(e)); }
    return 0;
}

int launch_gemm_tn_grouped(const TnJobTable& tab, int block_n, cudaStream_t st) {
    const int sms = device_sm_count();
    if (tab.n_jobs < 1 || tab.n_jobs > kTnMaxJobs || tab.total_items < 1) return set_error("gemm_tn (grouped): empty or oversized job table");
    if (block_n == 256) return launch_tn_grouped_cfg<256, 4>(tab, sms, st);
    if (block_n == 128) return launch_tn_grouped_cfg<128, 6>(tab, sms, st);
    if (block_n == 64)  return launch_tn_grouped_cfg<64, 8>(tab, sms, st);
    return set_error("gemm_tn (grouped): unsupported block_n %d", block_n);
}

template <typename T, int BLOCK_N, int STAGES, int CL, bool ORDERED = false>
static int launch_tn_cfg(const CUtensorMap& a, const CUtensorMap& b, const GemmTnParams& p, int num_sms, cudaStream_t st) {
    constexpr int SLAB = 128 / sizeof(T);
    constexpr int stage_bytes = (128 / SLAB + BLOCK_N / SLAB) * kTnBK * 128;
    constexpr int total = STAGES * stage_bytes + (2 * STAGES + 4) * 8 + 16 + 1024;
    static_assert(total <= 227 * 1024, "shared memory budget");
    auto kern = gemm_tn_kernel<T, BLOCK_N, STAGES, CL, ORDERED>;
    static bool attr_set = false;
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kTnThreads);
    cfg.dynamicSmemBytes = total;
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
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, total);
        if (e != cudaSuccess) return set_error("gemm_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        if (CL > 1) {
            cfg.gridDim = dim3((num_sms / CL) * CL);
            int n = 0;
            e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
            if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = num_sms / CL; }
            max_clusters = n;
        }
        attr_set = true;
    }
    const int batches = p.batch_mode ? p.H * p.N : 1;
    int grid;
    if (CL > 1) {
        const long long pairs = (long long)p.splits * (p.m_tiles / CL) * p.n_tiles * p.taps;
        const long long cl_cap = max_clusters < num_sms / CL ? max_clusters : num_sms / CL;
        grid = (int)(pairs < cl_cap ? pairs : cl_cap) * CL;
    } else {
        const long long items = (long long)p.splits * p.m_tiles * p.n_tiles * p.taps * batches;
        grid = (int)(items < num_sms ? items : num_sms);
    }
    cfg.gridDim = dim3(grid);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, b, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return set_error("gemm_tn launch: %s", cudaGetErrorString(e)); }
    return 0;
}

int launch_gemm_tn(int dtype, const CUtensorMap& a, const CUtensorMap& b, const GemmTnParams& p, int block_n, cudaStream_t st) {
    const int sms = device_sm_count();
    const bool pair = p.cluster == 2 && !p.batch_mode && p.m_tiles % 2 == 0 && dtype == 0 && block_n >= 128;
    const bool ordered = p.out_mode == 0 && p.splits > 1 && p.ws != nullptr;
    if (ordered) {          // opt-in (deterministic mode): unclustered kernels only
        if (dtype == 0) {
            if (block_n == 256) return launch_tn_cfg<__nv_bfloat16, 256, 4, 1, true>(a, b, p, sms, st);
            if (block_n == 128) return launch_tn_cfg<__nv_bfloat16, 128, 6, 1, true>(a, b, p, sms, st);
            if (block_n == 64)  return launch_tn_cfg<__nv_bfloat16, 64, 8, 1, true>(a, b, p, sms, st);
        } else {
            if (block_n == 128) return launch_tn_cfg<float, 128, 3, 1, true>(a, b, p, sms, st);
            if (block_n == 64)  return launch_tn_cfg<float, 64, 4, 1, true>(a, b, p, sms, st);
            if (block_n == 32)  return launch_tn_cfg<float, 32, 5, 1, true>(a, b, p, sms, st);
        }
        return set_error("gemm_tn: unsupported block_n %d for dtype %d", block_n, dtype);
    }
    if (dtype == 0) {
        if (pair) {
            if (block_n == 256) return launch_tn_cfg<__nv_bfloat16, 256, 4, 2>(a, b, p, sms, st);
            return launch_tn_cfg<__nv_bfloat16, 128, 6, 2>(a, b, p, sms, st);
        }
        if (block_n == 256) return launch_tn_cfg<__nv_bfloat16, 256, 4, 1>(a, b, p, sms, st);
        if (block_n == 128) return launch_tn_cfg<__nv_bfloat16, 128, 6, 1>(a, b, p, sms, st);
        if (block_n == 64)  return launch_tn_cfg<__nv_bfloat16, 64, 8, 1>(a, b, p, sms, st);
    } else {
        if (block_n == 128) return launch_tn_cfg<float, 128, 3, 1>(a, b, p, sms, st);
        if (block_n == 64)  return launch_tn_cfg<float, 64, 4, 1>(a, b, p, sms, st);
        if (block_n == 32)  return launch_tn_cfg<float, 32, 5, 1>(a, b, p, sms, st);
    }
    return set_error("gemm_tn: unsupported block_n %d for dtype %d", block_n, dtype);
}

}  // namespace b2
