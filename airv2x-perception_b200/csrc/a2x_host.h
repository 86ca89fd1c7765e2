// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
// The driver entry point for cuTensorMapEncodeTiled is resolved at run time (cudaGetDriverEntryPoint) so the
// library has no link-time dependency on libcuda and still loads on a CPU-only box (symbol-export tests).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace a2x {

void set_error(const char* fmt, ...);  // thread-local message, read through a2x_last_error()
extern unsigned long long g_launches;  // kernels launched by this library (a2x_launch_count)
#define A2X_LAUNCHED() (++a2x::g_launches)

#define A2X_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            a2x::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

#define A2X_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            a2x::set_error(__VA_ARGS__); \
            return 1;                   \
        }                               \
    } while (0)

// Encode a rank-`rank` fp32 tiled tensor map with 128-byte swizzle. dims/box are in elements (dim 0 innermost,
// contiguous); strides_bytes[i] is the byte stride of dim i+1 (rank-1 entries). Returns 0 on success.
// swizzle_atom32 != 0 selects CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (for MN-major tf32 UMMA operands).
// bf16 != 0 encodes a bfloat16 tensor (dims/box still in elements, strides in bytes).
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_atom32 = 0, int bf16 = 0);

}  // namespace a2x
