// Lift-Splat camera branch: the "lift" (depth distribution x image features, CamEncode.forward,
// opencood/models/sub_modules/lss_submodule.py:170-186) fused with the "splat" (LiftSplatShootEncoder.voxel_pooling,
// opencood/models/common_modules/airv2x_encoder.py:208-275). The reference materialises the [B,N,D,fH,fW,C] product
// (48 x the feature map), sorts its rows by voxel rank and pools them with a cumulative-sum trick; here one warp per
// image pixel keeps the pixel's C features in registers, walks the D depth bins and adds depth * feature straight into
// the BEV cell the frustum point falls in (vector atomics into an NHWC canvas, channel = z * C + c as the reference's
// `torch.cat(final.unbind(dim=2), 1)`). HBM-bound: reads depth + features + geometry once, writes only touched cells.
//
// The cell index reproduces `((geom - (bx - dx / 2)) / dx).long()` in fp32 (truncation toward zero, so points up to one
// cell below the grid fall into cell 0 like in the reference). Sums differ from the reference only by its own
// cumulative-sum cancellation error (both are compared with the float64 sum in tests/test_gpu_lss.py).
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"

namespace a2x {

struct LssGrid {
    float s[3], dx[3];   // s = bx - dx / 2 (fp32, computed by the host exactly like the reference tensor expression)
    int nx[3];
};

__device__ __forceinline__ int lss_cell(const float* __restrict__ g3, const LssGrid& gr, int b) {
    int c[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        c[j] = (int)__fdiv_rn(__fsub_rn(g3[j], gr.s[j]), gr.dx[j]);   // float -> int truncates toward zero (= .long())
        if (c[j] < 0 || c[j] >= gr.nx[j]) return -1;
    }
    return ((b * gr.nx[1] + c[1]) * gr.nx[0] + c[0]) * gr.nx[2] + c[2];   // NHWC canvas cell, z innermost
}

// warp = one image pixel (bn, h, w); half-warp = one depth bin at a time; lane q of a half-warp = channels 4q .. 4q+3
template <int C4>   // C / 4 <= 16
__global__ void __launch_bounds__(256) lift_splat_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ feat,
                                                             const float* __restrict__ geom, int BN, int N, int D, int HW,
                                                             LssGrid gr, float* __restrict__ bev, int* __restrict__ cells) {
    const int C = C4 * 4;
    const int lane = threadIdx.x & 31, half = lane >> 4, q = lane & 15;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp; pix < (long long)BN * HW; pix += nwarps) {
        const int bn = (int)(pix / HW), hw = (int)(pix - (long long)bn * HW);
        const int b = bn / N;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < C4) {
            const float* fp = feat + ((long long)bn * C + q * 4) * HW + hw;
            f = make_float4(fp[0], fp[HW], fp[2 * (long long)HW], fp[3 * (long long)HW]);
        }
        for (int d = half; d < D; d += 2) {
            const long long pt = ((long long)bn * D + d) * HW + hw;
            const int cell = lss_cell(geom + pt * 3, gr, b);
            if (q == 0) cells[pt] = cell;
            if (cell < 0 || q >= C4) continue;
            const float p = depth[pt];
            float4* dst = reinterpret_cast<float4*>(bev + (long long)cell * C + q * 4);
            atomicAdd(dst, make_float4(__fmul_rn(p, f.x), __fmul_rn(p, f.y), __fmul_rn(p, f.z), __fmul_rn(p, f.w)));
        }
    }
}

// backward: d depth[pt] = sum_c dbev[cell][c] * feat[c];  d feat[c] = sum_d depth[pt] * dbev[cell][c]   (gathers only)
template <int C4>
__global__ void __launch_bounds__(256) lift_splat_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ feat,
                                                             const int* __restrict__ cells, const float* __restrict__ dbev,
                                                             int BN, int D, int HW, float* __restrict__ ddepth,
                                                             float* __restrict__ dfeat) {
    const int C = C4 * 4;
    const int lane = threadIdx.x & 31, half = lane >> 4, q = lane & 15;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp; pix < (long long)BN * HW; pix += nwarps) {
        const int bn = (int)(pix / HW), hw = (int)(pix - (long long)bn * HW);
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f), acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < C4) {
            const float* fp = feat + ((long long)bn * C + q * 4) * HW + hw;
            f = make_float4(fp[0], fp[HW], fp[2 * (long long)HW], fp[3 * (long long)HW]);
        }
        for (int d = half; d < D + half; d += 2) {     // both halves run the same trip count (shuffles below)
            const bool live = d < D;
            const long long pt = ((long long)bn * D + (live ? d : 0)) * HW + hw;
            const int cell = live ? cells[pt] : -1;
            float dot = 0.f;
            if (cell >= 0 && q < C4) {
                const float4 g = *reinterpret_cast<const float4*>(dbev + (long long)cell * C + q * 4);
                const float p = depth[pt];
                dot = g.x * f.x + g.y * f.y + g.z * f.z + g.w * f.w;
                acc.x = fmaf(p, g.x, acc.x); acc.y = fmaf(p, g.y, acc.y); acc.z = fmaf(p, g.z, acc.z); acc.w = fmaf(p, g.w, acc.w);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);   // within the half-warp
            if (live && q == 0) ddepth[pt] = dot;
        }
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
        if (half == 0 && q < C4) {
            float* o = dfeat + ((long long)bn * C + q * 4) * HW + hw;
            o[0] = acc.x; o[HW] = acc.y; o[2 * (long long)HW] = acc.z; o[3 * (long long)HW] = acc.w;
        }
    }
}

static int lss_grid_dims(long long pixels) {
    long long b = (pixels * 32 + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace a2x

extern "C" {

int a2x_lift_splat_fwd(const float* depth, const float* feat, const float* geom, int B, int N, int D, int fH, int fW, int C,
                       const float* origin3, const float* dx3, const int* nx3, float* bev, int* cells_ws, a2x_stream_t stream) {
    A2X_REQUIRE(depth && feat && geom && origin3 && dx3 && nx3 && bev && cells_ws && B > 0 && N > 0 && D > 0 && fH > 0 && fW > 0,
                "lift_splat_fwd: bad args");
    A2X_REQUIRE(C % 4 == 0 && C > 0 && C <= 64, "lift_splat_fwd: C must be a multiple of 4, at most 64");
    a2x::LssGrid gr;
    for (int j = 0; j < 3; ++j) {
        gr.s[j] = origin3[j]; gr.dx[j] = dx3[j]; gr.nx[j] = nx3[j];
    }
    A2X_REQUIRE((long long)B * nx3[0] * nx3[1] * nx3[2] < (1ll << 31), "lift_splat_fwd: BEV grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(bev, 0, sizeof(float) * (size_t)B * nx3[0] * nx3[1] * nx3[2] * C, st));
    const int HW = fH * fW, BN = B * N;
    const int g = a2x::lss_grid_dims((long long)BN * HW);
#define A2X_LSS(C4) a2x::lift_splat_fwd_kernel<C4><<<g, 256, 0, st>>>(depth, feat, geom, BN, N, D, HW, gr, bev, cells_ws)
    switch (C / 4) {
        case 16: A2X_LSS(16); break;
        case 8: A2X_LSS(8); break;
        case 4: A2X_LSS(4); break;
        default: a2x::set_error("lift_splat_fwd: C = %d not in {16, 32, 64}", C); return 1;
    }
#undef A2X_LSS
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_lift_splat_bwd(const float* depth, const float* feat, const int* cells_ws, const float* dbev, int B, int N, int D,
                       int fH, int fW, int C, float* ddepth, float* dfeat, a2x_stream_t stream) {
    A2X_REQUIRE(depth && feat && cells_ws && dbev && ddepth && dfeat && B > 0 && N > 0 && D > 0 && fH > 0 && fW > 0,
                "lift_splat_bwd: bad args");
    A2X_REQUIRE(C % 4 == 0 && C > 0 && C <= 64, "lift_splat_bwd: C must be a multiple of 4, at most 64");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = fH * fW, BN = B * N;
    const int g = a2x::lss_grid_dims((long long)BN * HW);
#define A2X_LSS(C4) a2x::lift_splat_bwd_kernel<C4><<<g, 256, 0, st>>>(depth, feat, cells_ws, dbev, BN, D, HW, ddepth, dfeat)
    switch (C / 4) {
        case 16: A2X_LSS(16); break;
        case 8: A2X_LSS(8); break;
        case 4: A2X_LSS(4); break;
        default: a2x::set_error("lift_splat_bwd: C = %d not in {16, 32, 64}", C); return 1;
    }
#undef A2X_LSS
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
