// Error reporting, debug knobs and TMA tensor-map encoding shared by every C-ABI entry point.
#include <stdarg.h>

#include "../../include/airv2x_b200.h"
#include "a2x_host.h"

namespace a2x {

static thread_local char g_err[1024] = "";
int g_debug[16] = {0};
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_atom32, int bf16) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return 2;
    cuuint64_t gdims[5], gstr[4];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i < rank - 1) gstr[i] = strides_bytes[i];
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
        set_error("tensor map base %p not 16-byte aligned", base);
        return 1;
    }
    CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u] "
                  "strides=[%llu,%llu,%llu,%llu]",
                  (int)r, rank, (unsigned long long)gdims[0], (unsigned long long)(rank > 1 ? gdims[1] : 0),
                  (unsigned long long)(rank > 2 ? gdims[2] : 0), (unsigned long long)(rank > 3 ? gdims[3] : 0),
                  (unsigned long long)(rank > 4 ? gdims[4] : 0), gbox[0], rank > 1 ? gbox[1] : 0, rank > 2 ? gbox[2] : 0,
                  rank > 3 ? gbox[3] : 0, rank > 4 ? gbox[4] : 0, (unsigned long long)gstr[0],
                  (unsigned long long)(rank > 2 ? gstr[1] : 0), (unsigned long long)(rank > 3 ? gstr[2] : 0),
                  (unsigned long long)(rank > 4 ? gstr[3] : 0));
        return 2;
    }
    return 0;
}

}  // namespace a2x

extern "C" {

const char* a2x_last_error(void) { return a2x::g_err; }

int a2x_version(void) { return 100; }

void a2x_debug_set(int key, int value) {
    if (key >= 0 && key < 16) a2x::g_debug[key] = value;
}

unsigned long long a2x_launch_count(void) { return a2x::g_launches; }

int a2x_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    A2X_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    A2X_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

}  // extern "C"
