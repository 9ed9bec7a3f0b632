// Host-side helpers: thread-local error string, device properties, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

int set_error(const char* fmt, ...);          // stores a thread-local message, returns a nonzero code
const char* last_error();
int device_sm_count();

// 4-D tiled tensor map with SWIZZLE_128B and zero out-of-bounds fill.
//  elem_bytes: 2 (bf16) or 4 (fp32). dims[0] is the contiguous dim; strides_bytes[i] is the stride of dims[i+1].
//  atom32: use SWIZZLE_128B_ATOM_32B (needed by MN-major tf32 operands of tcgen05.mma).
int make_tmap_4d(CUtensorMap* out, const void* base, int elem_bytes, const uint64_t dims[4],
                 const uint64_t strides_bytes[3], const uint32_t box[4], bool atom32 = false);

// 5-D variant for MN-major operands (gemm_tn.cu): dims = (elements of one 128-byte slab | d1 | d2 | d3 | slab index), the slab
// dimension having a stride of 128 bytes, so that a box of `box[4]` slabs lands in shared memory as consecutive
// [rows][128 B] slabs -- the layout the MN-major UMMA descriptor walks with its leading-byte offset -- from ONE TMA instruction.
int make_tmap_5d_slabs(CUtensorMap* out, const void* base, int elem_bytes, const uint64_t dims[4],
                       const uint64_t strides_bytes[3], const uint32_t box[4], uint32_t slabs_per_box, bool atom32 = false);

// Generic tiled map of rank 3..5 (SWIZZLE_128B, inner box = 128 bytes): dims / box have `rank` entries, strides rank - 1.
int make_tmap_nd(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box);

// Programmatic dependent launch (PDL): every kernel of this library begins with griddepcontrol.launch_dependents and
// executes griddepcontrol.wait before it touches global memory, so the next kernel's CTAs may be scheduled -- and run their
// prologue (barrier init, TMEM allocation, descriptor prefetch) -- while the previous grid drains, instead of paying the
// full launch + ramp latency between ~1.3 k back-to-back launches per step.  The attribute survives CUDA-graph capture
// (programmatic edges).  Measured on B200 inside CUDA-graph replay: a win of 4 % for small workloads (64x64 batch 8, ~1300
// launches of ~15 us) and a loss of 1 % for large ones (128x128 batch 32), so the host enables it per captured step by
// workload size (b200/graph.py; b2_set_option("pdl", 0|1)); SDM_B200_PDL=0|1 forces it.
bool pdl_enabled();

// Deterministic mode (SDM_B200_DETERMINISTIC=1 or b2_set_deterministic): no floating-point atomics and no batch-size dependent
// tiling on the forward path -- GroupNorm statistics come from the fixed-order statistics kernel instead of the conv epilogue's
// atomics, split-K is off for the NT kernel -- so an image's result does not depend on what else is in the batch.
bool deterministic_mode();
void set_deterministic_mode(int on);

// Kernel-selection switches (A/B measurements and tests): initialised from SDM_B200_<NAME> (upper case) on first use, changed at
// run time through b2_set_option.  "halo": halo-tile 3x3 convs; "swap_ab": swapped-operand convs for <= 128 output channels;
// "sm_limit" (SDM_B200_SM_LIMIT): SMs the persistent grids may occupy (0 = all) -- data-parallel training can leave a few SMs
// to NCCL's copy/reduce CTAs, which otherwise queue behind 148-CTA persistent GEMMs.
int option(const char* name, int default_value);
int set_option(const char* name, int value);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace b2

#define B2_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (void)b2::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), __VA_ARGS__)
