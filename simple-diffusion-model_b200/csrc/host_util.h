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

}  // namespace b2
