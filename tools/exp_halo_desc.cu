// Experiment (B200 only): can a SWIZZLE_128B K-major UMMA operand start at an arbitrary 128-byte ROW offset inside a
// TMA-written tile?  A halo tile reused across the 9 taps of a 3x3 conv needs exactly that (a tap = a row shift).
// For every shift s the A operand is rows [s, s+128) of one TMA box of 272 rows; the descriptor's start address is
// base + s*128 and its "matrix base offset" field (bits 49-51) is tried as 0 and as (start >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Isimple-diffusion-model_b200/csrc -Iinclude tools/exp_halo_desc.cu -o build/exp_halo_desc -lcuda
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>

using namespace b2;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int ROWS = 256, ROWS2 = 16;       // box of ROWS + a second box of ROWS2 rows right behind it (272 rows in total)
constexpr int NB = 64;                      // B rows (N of the MMA)

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = umma_desc_sw128(saddr, 16, sbo);
    d |= static_cast<uint64_t>(base_off & 7) << 49;
    return d;
}

__global__ void __launch_bounds__(128, 1)
exp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
           float* out, int shift, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = smem;                                  // 272 rows x 128 B
    uint8_t* b_s = smem + 36 * 1024;                      // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint64_t* mbar = bar + 1;
    uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<64>(holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, (ROWS + ROWS2 + NB) * 128u);
        tma_load_4d(a_s, &tmA, bar, 0, 0, 0, 0);
        tma_load_4d(a_s + ROWS * 128, &tmA2, bar, 0, ROWS, 0, 0);
        tma_load_4d(b_s, &tmB, bar, 0, 0, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc(1u, 128, NB, 0, 0);
        const uint32_t a_addr = smem_u32(a_s) + shift * 128;
        const uint32_t b_addr = smem_u32(b_s);
        const uint32_t bo = mode == 0 ? 0u : ((a_addr >> 7) & 7u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_ss<false>(tmem, desc_bo(a_addr + k * 32, 1024, bo), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc, k ? 1u : 0u);
        umma_commit(mbar);
    }
    mbar_wait(mbar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < NB; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * NB + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_enc() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}
static void make_map(CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows) {
    cuuint64_t gd[4] = {64, rows, 1, 1}, gs[3] = {128, 128 * rows, 128 * rows};
    cuuint32_t bx[4] = {64, box_rows, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = get_enc()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(3); }
}

int main() {
    const int TOT = ROWS + ROWS2;
    std::vector<__nv_bfloat16> ha(TOT * 64), hb(NB * 64);
    std::vector<float> fa(TOT * 64), fb(NB * 64);
    uint32_t s = 1;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (int i = 0; i < TOT * 64; ++i) { ha[i] = __float2bfloat16(rnd()); fa[i] = __bfloat162float(ha[i]); }
    for (int i = 0; i < NB * 64; ++i) { hb[i] = __float2bfloat16(rnd()); fb[i] = __bfloat162float(hb[i]); }
    void *da, *db; float* dout;
    CK(cudaMalloc(&da, TOT * 128)); CK(cudaMalloc(&db, NB * 128)); CK(cudaMalloc(&dout, 128 * NB * 4));
    CK(cudaMemcpy(da, ha.data(), TOT * 128, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), NB * 128, cudaMemcpyHostToDevice));
    CUtensorMap ta, ta2, tb;
    make_map(&ta, da, TOT, ROWS); make_map(&ta2, da, TOT, ROWS2); make_map(&tb, db, NB, NB);
    CK(cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int shifts[] = {0, 8, 1, 2, 7, 9, 33, 65, 66, 67, 130, 133, 134};
    for (int mode = 0; mode < 2; ++mode)
        for (int sh : shifts) {
            CK(cudaMemset(dout, 0, 128 * NB * 4));
            exp_kernel<<<1, 128, 64 * 1024>>>(ta, ta2, tb, dout, sh, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d shift %3d: CUDA error %s\n", mode, sh, cudaGetErrorString(e)); return 1; }
            std::vector<float> ho(128 * NB);
            CK(cudaMemcpy(ho.data(), dout, 128 * NB * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < NB; ++n) {
                double acc = 0;
                for (int k = 0; k < 64; ++k) acc += (double)fa[(m + sh) * 64 + k] * fb[n * 64 + k];
                double d = fabs(acc - ho[m * NB + n]); if (d > maxerr) maxerr = d;
            }
            printf("base_offset %-14s shift %3d rows: max_err %.3e  %s\n", mode == 0 ? "0" : "(addr>>7)&7", sh, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
