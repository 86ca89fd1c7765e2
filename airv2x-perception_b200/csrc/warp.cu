// Ego-warp: affine bilinear resampling of NHWC BEV feature maps (forward and backward), one fused gather kernel
// instead of F.affine_grid + F.grid_sample.
//
// Reference: warp_affine_simple (opencood/models/common_modules/torch_transformation_utils.py:327-334) =
//   grid = F.affine_grid(M, [B, C, Ho, Wo], align_corners); out = F.grid_sample(src, grid, bilinear, zeros padding,
//   align_corners), with M the normalised 2x3 matrices built at where2comm_attn.py:293-307 /
//   utils/transformation_utils.py:396-422 (host side: a2x host module `warp.normalize_pairwise`).
//   base grid (align_corners = False): x_i = (2 i + 1) / Wo - 1;  source index ix = ((gx + 1) * Wi - 1) / 2.
//   base grid (align_corners = True):  x_i = 2 i / (Wo - 1) - 1;  source index ix = (gx + 1) / 2 * (Wi - 1).
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"

namespace a2x {

struct WarpGeom {
    int n, hi, wi, ho, wo, c, src_cs, dst_cs, align, nearest;
};

__device__ __forceinline__ void warp_coords(const float* __restrict__ th, const WarpGeom& g, int oy, int ox, float& ix,
                                            float& iy) {
    float bx, by;
    if (g.align) {
        bx = g.wo > 1 ? 2.f * ox / (g.wo - 1) - 1.f : 0.f;
        by = g.ho > 1 ? 2.f * oy / (g.ho - 1) - 1.f : 0.f;
    } else {
        bx = (2.f * ox + 1.f) / g.wo - 1.f;
        by = (2.f * oy + 1.f) / g.ho - 1.f;
    }
    const float gx = th[0] * bx + th[1] * by + th[2];
    const float gy = th[3] * bx + th[4] * by + th[5];
    if (g.align) {
        ix = (gx + 1.f) * 0.5f * (g.wi - 1);
        iy = (gy + 1.f) * 0.5f * (g.hi - 1);
    } else {
        ix = ((gx + 1.f) * g.wi - 1.f) * 0.5f;
        iy = ((gy + 1.f) * g.hi - 1.f) * 0.5f;
    }
}

// one thread per (output pixel, channel quad)
__global__ void __launch_bounds__(256) warp_affine_fwd_kernel(const float* __restrict__ src,
                                                              const float* __restrict__ theta, SplitOut dst,
                                                              const WarpGeom g) {
    const int q = g.c >> 2;
    const long long total = (long long)g.n * g.ho * g.wo * q;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % q);
        long long pix = i / q;
        const int ox = (int)(pix % g.wo);
        const int oy = (int)((pix / g.wo) % g.ho);
        const int img = (int)(pix / ((long long)g.wo * g.ho));
        float ix, iy;
        warp_coords(theta + img * 6, g, oy, ox, ix, iy);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* base = src + (long long)img * g.hi * g.wi * g.src_cs + cq * 4;
        if (g.nearest) {
            const int x0 = (int)nearbyintf(ix), y0 = (int)nearbyintf(iy);
            if (x0 >= 0 && x0 < g.wi && y0 >= 0 && y0 < g.hi)
                acc = *reinterpret_cast<const float4*>(base + ((long long)y0 * g.wi + x0) * g.src_cs);
        } else {
            const float fx = floorf(ix), fy = floorf(iy);
            const int x0 = (int)fx, y0 = (int)fy;
            const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
            const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
                if (xx >= 0 && xx < g.wi && yy >= 0 && yy < g.hi) {
                    const float4 v = *reinterpret_cast<const float4*>(base + ((long long)yy * g.wi + xx) * g.src_cs);
                    acc.x = fmaf(wgt[t], v.x, acc.x);
                    acc.y = fmaf(wgt[t], v.y, acc.y);
                    acc.z = fmaf(wgt[t], v.z, acc.z);
                    acc.w = fmaf(wgt[t], v.w, acc.w);
                }
            }
        }
        store_split4(dst, pix * g.dst_cs + cq * 4, acc);
    }
}

// d(src) += bilinear^T d(out): scatter with fp32 atomics (dsrc zeroed by the caller)
__global__ void __launch_bounds__(256) warp_affine_bwd_kernel(const float* __restrict__ dout,
                                                              const float* __restrict__ theta,
                                                              float* __restrict__ dsrc, const WarpGeom g) {
    const int q = g.c >> 2;
    const long long total = (long long)g.n * g.ho * g.wo * q;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % q);
        long long pix = i / q;
        const int ox = (int)(pix % g.wo);
        const int oy = (int)((pix / g.wo) % g.ho);
        const int img = (int)(pix / ((long long)g.wo * g.ho));
        float ix, iy;
        warp_coords(theta + img * 6, g, oy, ox, ix, iy);
        const float4 d = *reinterpret_cast<const float4*>(dout + pix * g.dst_cs + cq * 4);
        float* base = dsrc + (long long)img * g.hi * g.wi * g.src_cs + cq * 4;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
        const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
            if (xx >= 0 && xx < g.wi && yy >= 0 && yy < g.hi) {
                float* o = base + ((long long)yy * g.wi + xx) * g.src_cs;
                atomicAdd(o, wgt[t] * d.x);
                atomicAdd(o + 1, wgt[t] * d.y);
                atomicAdd(o + 2, wgt[t] * d.z);
                atomicAdd(o + 3, wgt[t] * d.w);
            }
        }
    }
}

// ROI mask = nearest-neighbour warp of an all-ones map (get_rotated_roi, torch_transformation_utils.py:81-113) times the
// per-agent validity flag: out[a][p] = valid[a] && the source pixel of p falls inside the map
__global__ void roi_mask_kernel(const float* __restrict__ theta, const int* __restrict__ valid, float* __restrict__ out,
                                const WarpGeom g) {
    const long long total = (long long)g.n * g.ho * g.wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % g.wo);
        const int oy = (int)((i / g.wo) % g.ho);
        const int img = (int)(i / ((long long)g.wo * g.ho));
        float ix, iy;
        warp_coords(theta + img * 6, g, oy, ox, ix, iy);
        const int x0 = (int)nearbyintf(ix), y0 = (int)nearbyintf(iy);
        const bool in = x0 >= 0 && x0 < g.wi && y0 >= 0 && y0 < g.hi;
        out[i] = (in && (valid == nullptr || valid[img] != 0)) ? 1.f : 0.f;
    }
}

static int warp_grid(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_warp_affine_fwd(const float* src, int src_cs, const float* theta, int n, int hi, int wi, int c, int ho, int wo,
                        int align_corners, int nearest, const a2x_output* dst, a2x_stream_t stream) {
    A2X_REQUIRE(src && theta && dst && dst->hi && n > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0 && c % 4 == 0 &&
                    src_cs >= c && dst->cs >= c,
                "warp_affine_fwd: bad args (channels must be a multiple of 4)");
    WarpGeom g{n, hi, wi, ho, wo, c, src_cs, dst->cs, align_corners, nearest};
    SplitOut o;
    o.hi = dst->hi;
    o.b16 = (__nv_bfloat16*)dst->b16;
    o.ps = dst->b16_plane;
    warp_affine_fwd_kernel<<<warp_grid((long long)n * ho * wo * (c / 4)), 256, 0, (cudaStream_t)stream>>>(src, theta, o, g);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_roi_mask(const float* theta, const int* valid, int n, int h, int w, int align_corners, float* out,
                 a2x_stream_t stream) {
    A2X_REQUIRE(theta && out && n > 0 && h > 0 && w > 0, "roi_mask: bad args");
    WarpGeom g{n, h, w, h, w, 4, 4, 4, align_corners, 1};
    roi_mask_kernel<<<warp_grid((long long)n * h * w), 256, 0, (cudaStream_t)stream>>>(theta, valid, out, g);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_warp_affine_bwd(const float* dout, int dout_cs, const float* theta, int n, int hi, int wi, int c, int ho, int wo,
                        int align_corners, float* dsrc, int dsrc_cs, a2x_stream_t stream) {
    A2X_REQUIRE(dout && theta && dsrc && n > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0 && c > 0 && c % 4 == 0 &&
                    dsrc_cs >= c && dout_cs >= c,
                "warp_affine_bwd: bad args (channels must be a multiple of 4)");
    WarpGeom g{n, hi, wi, ho, wo, c, dsrc_cs, dout_cs, align_corners, 0};
    warp_affine_bwd_kernel<<<warp_grid((long long)n * ho * wo * (c / 4)), 256, 0, (cudaStream_t)stream>>>(dout, theta,
                                                                                                        dsrc, g);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
