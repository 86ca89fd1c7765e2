// Counter-based dropout masks shared by the dropout kernels (dropout.cu) and the tap-GEMM epilogue (tapgemm.cuh):
//     keep(seed, s, e) = u16(philox4x32_10(key = seed, counter = (e / 8, s))[e % 8]) >= round(p * 65536)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace a2x {

__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 ctr) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

struct DropArgs {
    unsigned long long seed;
    uint32_t site, thresh;   // thresh = round(p * 65536); 0 = dropout off
    float scale;             // 1 / (1 - p)
};

// multipliers (0 or scale) of elements 8 g .. 8 g + 7
__device__ __forceinline__ void drop_mult8(const DropArgs& d, long long g, float (&m)[8]) {
    const uint4 r = philox4x32_10(make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)),
                                  make_uint4((uint32_t)g, (uint32_t)((unsigned long long)g >> 32), d.site, 0x0A2Du));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = ((w[k >> 1] >> (16 * (k & 1))) & 0xffffu) >= d.thresh ? d.scale : 0.f;
}

}  // namespace a2x
