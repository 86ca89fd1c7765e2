// nn.Dropout of the transformer fusion networks in train mode (cobevt_modules/base_transformer.py:27-56,
// swap_fusion_modules.py:43, v2xvit_modules/base_transformer.py:17-46, hmsa.py:18,155, mswin.py:47) with a counter-based
// generator: the keep flag of element e of dropout site s in step `seed` is a pure function
//     keep(seed, s, e) = u16(philox4x32_10(key = seed, counter = (e / 8, s))[e % 8]) >= round(p * 65536)
// so the forward kernels and the backward kernels regenerate the SAME mask and nothing is stored (the reference keeps a
// byte mask per site for autograd). One Philox call yields the flags of 8 consecutive elements. Kept values are scaled by
// 1 / (1 - p). a2x_dropout_mask exports the mask a site uses, so that a test can feed identical masks to the oracle.
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"
#include "philox.cuh"

namespace a2x {

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_df(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// MODE 0: out = (res ? res : 0) + y * m        (dropout, optionally fused with the residual add; res may alias out.hi)
// MODE 1: out = gelu(y) * m                    (FeedForward: Linear -> GELU -> Dropout)
// MODE 2: out = y * m * gelu'(aux)             (its backward; aux = the pre-activation)
template <int MODE>
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ y, const float* __restrict__ aux,
                                                      const float* res, long long n8, DropArgs d, SplitOut out) {
    // U independent groups of 8 elements per thread and iteration: all loads of the iteration are issued before the first
    // use (the one-group loop was latency-bound: 2.3x off the HBM rate with two loads in flight per thread)
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; g0 < n8; g0 += stride * U) {
        float4 ya[U], yb[U], xa[U], xb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long g = g0 + u * stride;
            if (g < n8) {
                ya[u] = reinterpret_cast<const float4*>(y)[2 * g];
                yb[u] = reinterpret_cast<const float4*>(y)[2 * g + 1];
                if (MODE == 2) {
                    xa[u] = reinterpret_cast<const float4*>(aux)[2 * g];
                    xb[u] = reinterpret_cast<const float4*>(aux)[2 * g + 1];
                } else if (MODE == 0 && res != nullptr) {
                    xa[u] = reinterpret_cast<const float4*>(res)[2 * g];
                    xb[u] = reinterpret_cast<const float4*>(res)[2 * g + 1];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long g = g0 + u * stride;
            if (g >= n8) break;
            float m[8];
            if (d.thresh != 0) {
                drop_mult8(d, g, m);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) m[k] = 1.f;
            }
            float v[8] = {ya[u].x, ya[u].y, ya[u].z, ya[u].w, yb[u].x, yb[u].y, yb[u].z, yb[u].w};
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = gelu_f(v[k]) * m[k];
            } else if (MODE == 2) {
                const float x[8] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w, xb[u].x, xb[u].y, xb[u].z, xb[u].w};
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = v[k] * m[k] * gelu_df(x[k]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] *= m[k];
                if (res != nullptr) {
                    v[0] += xa[u].x; v[1] += xa[u].y; v[2] += xa[u].z; v[3] += xa[u].w;
                    v[4] += xb[u].x; v[5] += xb[u].y; v[6] += xb[u].z; v[7] += xb[u].w;
                }
            }
            store_split4(out, 8 * g, make_float4(v[0], v[1], v[2], v[3]));
            store_split4(out, 8 * g + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
    }
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(long long n8, DropArgs d, unsigned char* __restrict__ mask) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n8; g += (long long)gridDim.x * blockDim.x) {
        float m[8];
        drop_mult8(d, g, m);
#pragma unroll
        for (int k = 0; k < 8; ++k) mask[8 * g + k] = m[k] != 0.f ? 1 : 0;
    }
}

static SplitOut tr_split(const a2x_output* o) {
    SplitOut s;
    s.hi = o->hi;
    s.b16 = (__nv_bfloat16*)o->b16;
    s.ps = o->b16_plane;
    return s;
}
static DropArgs make_drop(unsigned long long seed, unsigned int site, float p) {
    DropArgs d;
    d.seed = seed;
    d.site = site;
    d.thresh = p > 0.f ? (uint32_t)(p * 65536.0f + 0.5f) : 0u;
    d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
    return d;
}
static int grid8(long long n8) {
    long long b = (n8 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_dropout_apply(const float* y, const float* residual, long long n, unsigned long long seed, unsigned int site, float p,
                      const a2x_output* out, a2x_stream_t stream) {
    A2X_REQUIRE(y && out && (out->hi || out->b16) && n > 0 && n % 8 == 0 && p >= 0.f && p < 1.f,
                "dropout_apply: bad args (n must be a multiple of 8, 0 <= p < 1)");
    dropout_kernel<0><<<grid8(n / 8), 256, 0, (cudaStream_t)stream>>>(y, nullptr, residual, n / 8, make_drop(seed, site, p), tr_split(out));
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_gelu_dropout_fwd(const float* x, long long n, unsigned long long seed, unsigned int site, float p, const a2x_output* y,
                         a2x_stream_t stream) {
    A2X_REQUIRE(x && y && (y->hi || y->b16) && n > 0 && n % 8 == 0 && p >= 0.f && p < 1.f, "gelu_dropout_fwd: bad args");
    dropout_kernel<1><<<grid8(n / 8), 256, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, n / 8, make_drop(seed, site, p), tr_split(y));
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_gelu_dropout_bwd(const float* dy, const float* x, long long n, unsigned long long seed, unsigned int site, float p,
                         const a2x_output* dx, a2x_stream_t stream) {
    A2X_REQUIRE(dy && x && dx && (dx->hi || dx->b16) && n > 0 && n % 8 == 0 && p >= 0.f && p < 1.f, "gelu_dropout_bwd: bad args");
    dropout_kernel<2><<<grid8(n / 8), 256, 0, (cudaStream_t)stream>>>(dy, x, nullptr, n / 8, make_drop(seed, site, p), tr_split(dx));
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_dropout_mask(long long n, unsigned long long seed, unsigned int site, float p, unsigned char* mask, a2x_stream_t stream) {
    A2X_REQUIRE(mask && n > 0 && n % 8 == 0 && p > 0.f && p < 1.f, "dropout_mask: bad args");
    dropout_mask_kernel<<<grid8(n / 8), 256, 0, (cudaStream_t)stream>>>(n / 8, make_drop(seed, site, p), mask);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
