#include "host_util.h"
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace b2 {

static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
const char* last_error() { return g_err; }

static int g_sm_limit = -1;      // "sm_limit" option: persistent grids use at most this many SMs (0 = all)
int device_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (sms[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    if (g_sm_limit < 0) { const char* e = getenv("SDM_B200_SM_LIMIT"); g_sm_limit = e ? atoi(e) : 0; }
    return (g_sm_limit > 0 && g_sm_limit < sms[dev]) ? g_sm_limit : sms[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_4d(CUtensorMap* out, const void* base, int elem_bytes, const uint64_t dims[4],
                 const uint64_t strides_bytes[3], const uint32_t box[4], bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error("TMA base pointer not 16-byte aligned");
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < 3; ++i) {
        gs[i] = strides_bytes[i];
        if (gs[i] % 16 != 0) return set_error("TMA stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
    }
    if (box[0] * elem_bytes != 128) return set_error("TMA inner box must be 128 bytes");
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error("cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu box %u %u %u %u", (int)r,
                         (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                         (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
    return 0;
}

int make_tmap_5d_slabs(CUtensorMap* out, const void* base, int elem_bytes, const uint64_t dims[4],
                       const uint64_t strides_bytes[3], const uint32_t box[4], uint32_t slabs_per_box, bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error("TMA base pointer not 16-byte aligned");
    const uint64_t slab = 128 / elem_bytes;
    if (dims[0] % slab) return set_error("TMA slab map: %llu channels are not whole 128-byte slabs", (unsigned long long)dims[0]);
    cuuint64_t gd[5] = {slab, dims[1], dims[2], dims[3], dims[0] / slab};
    cuuint64_t gs[4] = {strides_bytes[0], strides_bytes[1], strides_bytes[2], 128};
    cuuint32_t bx[5] = {(cuuint32_t)slab, box[1], box[2], box[3], slabs_per_box};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 3; ++i)
        if (gs[i] % 16 != 0) return set_error("TMA stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, 5, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error("cuTensorMapEncodeTiled (5-D slabs) failed (%d): dims %llu %llu %llu %llu box %u %u %u x %u slabs", (int)r,
                         (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                         (unsigned long long)dims[3], box[1], box[2], box[3], slabs_per_box);
    return 0;
}

int make_tmap_nd(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if (rank < 3 || rank > 5) return set_error("make_tmap_nd: rank %d", rank);
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error("TMA base pointer not 16-byte aligned");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) {
        gs[i] = strides_bytes[i];
        if (gs[i] % 16 != 0) return set_error("TMA stride %d (%llu B) not a multiple of 16", i, (unsigned long long)gs[i]);
    }
    if (box[0] * elem_bytes != 128) return set_error("TMA inner box must be 128 bytes");
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r);
    return 0;
}

static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("SDM_B200_PDL");
        g_pdl = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pdl == 1;
}

static int g_det = -1;
bool deterministic_mode() {
    if (g_det < 0) {
        const char* e = getenv("SDM_B200_DETERMINISTIC");
        g_det = (e && e[0] == '1') ? 1 : 0;
    }
    return g_det == 1;
}
void set_deterministic_mode(int on) { g_det = on ? 1 : 0; }

struct Opt { const char* name; const char* env; int value; bool set; };
static Opt g_opts[] = {{"halo", "SDM_B200_HALO", 0, false}, {"swap_ab", "SDM_B200_SWAP_AB", 0, false},
                       {"tn_box5", "SDM_B200_TN_BOX5", 0, false}, {"l2_prefetch", "SDM_B200_L2_PREFETCH", 0, false}};
int option(const char* name, int default_value) {
    for (auto& o : g_opts) {
        if (strcmp(o.name, name)) continue;
        if (!o.set) { const char* e = getenv(o.env); o.value = e ? atoi(e) : default_value; o.set = true; }
        return o.value;
    }
    return default_value;
}
int set_option(const char* name, int value) {
    if (!strcmp(name, "sm_limit")) { g_sm_limit = value > 0 ? value : 0; return 0; }
    if (!strcmp(name, "pdl")) { g_pdl = value ? 1 : 0; return 0; }
    for (auto& o : g_opts) if (!strcmp(o.name, name)) { o.value = value; o.set = true; return 0; }
    return set_error("b2_set_option: unknown option '%s'", name);
}

}  // namespace b2

// Zero-fill of a device range on `stream` (the flat gradient buffer at the start of every backward pass): a memset node, no kernel.
extern "C" int b2_set_option(const char* name, int value) { return b2::set_option(name, value); }

extern "C" int b2_zero(void* ptr, long long bytes, void* stream) {
    if (!ptr || bytes <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return b2::set_error("b2_zero: %s", cudaGetErrorString(e));
    return 0;
}
