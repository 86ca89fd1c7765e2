// HBM-bound elementwise / per-channel-reduction kernels around the tensor-core contractions:
// BatchNorm (train-mode batch statistics, eval-mode affine), ReLU, their backward passes, mask multiply,
// operand splitting for the 3-pass bf16 GEMMs. All tensors are fp32 NHWC with an explicit pixel stride.
//
// Reference semantics: nn.BatchNorm2d(eps=1e-3, momentum=0.01) + nn.ReLU in
// opencood/models/common_modules/base_bev_backbone.py:52-66, :82-90; bias+ReLU in downsample_conv.py:18-32.
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"

namespace a2x {

static __host__ SplitOut to_split(const a2x_output* o) {
    SplitOut r;
    r.hi = o->hi;
    r.b16 = (__nv_bfloat16*)o->b16;
    r.ps = o->b16_plane;
    return r;
}

// ---------------------------------------------------------------------------------------------- split
__global__ void split_kernel(const float* __restrict__ x, long long n4, SplitOut o) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        store_split4(o, 4 * i, reinterpret_cast<const float4*>(x)[i]);
}

// ---------------------------------------------------------------------------------------------- channel stats
// sums[c] += sum_p x[p][c] ; sums[C + c] += sum_p x[p][c]^2     (double accumulators)
// optional second operand: MODE 1 computes g = dy * (z*scale+shift > 0), zhat = (z-mean)*invstd and accumulates
// sums[c] += g, sums[C+c] += g*zhat  (BatchNorm+ReLU backward reductions)
template <int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(const float* __restrict__ x, int x_cs,
                                                             const float* __restrict__ z, int z_cs,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, long long npix, int C,
                                                             double* __restrict__ sums, int replicas) {
    extern __shared__ float red[];  // [2][256][4]
    const int q = C >> 2;                  // float4 groups per pixel
    const int cq = threadIdx.x % q;        // my channel quad
    const int prow = threadIdx.x / q;      // my pixel lane within the block
    const int ppb = blockDim.x / q;        // pixels per block iteration
    float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    float sc[4], sh[4], mu[4], is[4];
    if (MODE == 1 && prow < ppb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[e] = scale[cq * 4 + e];
            sh[e] = shift[cq * 4 + e];
            mu[e] = mean[cq * 4 + e];
            is[e] = invstd[cq * 4 + e];
        }
    }
    if (prow < ppb) {
        constexpr int U = MODE == 1 ? 4 : 8;  // pixels in flight per thread (MODE 1 loads two tensors)
        const long long pstride = (long long)gridDim.x * ppb;
        for (long long p0 = (long long)blockIdx.x * ppb + prow; p0 < npix; p0 += pstride * U) {
            float4 xv[U], zv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long p = p0 + u * pstride;
                if (p < npix) {
                    xv[u] = *reinterpret_cast<const float4*>(x + p * x_cs + cq * 4);
                    if (MODE == 1) zv[u] = *reinterpret_cast<const float4*>(z + p * z_cs + cq * 4);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (p0 + u * pstride >= npix) break;
                const float a[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                if (MODE == 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s0[e] += a[e];
                        s1[e] += a[e] * a[e];
                    }
                } else {
                    const float b[4] = {zv[u].x, zv[u].y, zv[u].z, zv[u].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float g = (b[e] * sc[e] + sh[e] > 0.f) ? a[e] : 0.f;
                        s0[e] += g;
                        s1[e] += g * (b[e] - mu[e]) * is[e];
                    }
                }
            }
        }
    }
    float* r0 = red;
    float* r1 = red + blockDim.x * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        r0[threadIdx.x * 4 + e] = s0[e];
        r1[threadIdx.x * 4 + e] = s1[e];
    }
    __syncthreads();
    if (threadIdx.x < q) {
        double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
        for (int r = 0; r < ppb; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                d0[e] += r0[(r * q + threadIdx.x) * 4 + e];
                d1[e] += r1[(r * q + threadIdx.x) * 4 + e];
            }
        }
        // `replicas` copies of the accumulators ([replica][2C]) spread the same-address atomics of the many blocks; the
        // consumer adds the copies up
        double* dst = sums + (long long)(blockIdx.x % replicas) * 2 * C;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            atomicAdd(&dst[threadIdx.x * 4 + e], d0[e]);
            atomicAdd(&dst[C + threadIdx.x * 4 + e], d1[e]);
        }
    }
    if (replicas > 1) {
        // the last block to finish folds the copies into copy 0 (ticket counter behind the copies, zeroed with them)
        __shared__ bool last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + (long long)replicas * 2 * C);
            last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (last) {
            __threadfence();
            for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
                double t = 0.0;
                for (int r = 0; r < replicas; ++r) t += __ldcg(&sums[(long long)r * 2 * C + i]);
                sums[i] = t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- BN finalize
// batch stats -> (scale, shift, mean, invstd); running stats updated `n_updates` times (the reference evaluates the
// backbone several times per step on identical data: airv2x_where2com.py:119,124, where2comm_fuse.py:218).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int n_updates,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int C,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * invstd;
    shift[c] = b - (float)m * g * invstd;
    if (mean_out) mean_out[c] = (float)m;
    if (invstd_out) invstd_out[c] = invstd;
    if (running_mean != nullptr && n_updates > 0) {
        const float unbiased = (float)(count > 1 ? var * count / (count - 1) : var);
        float rm = running_mean[c], rv = running_var[c];
        for (int i = 0; i < n_updates; ++i) {
            rm = (1.f - momentum) * rm + momentum * (float)m;
            rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        running_mean[c] = rm;
        running_var[c] = rv;
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float invstd = 1.f / sqrtf(rv[c] + eps);
    const float s = gamma[c] * invstd;
    scale[c] = s;
    shift[c] = beta[c] - rm[c] * s;
}

// dgamma = sum g*zhat, dbeta = sum g
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float dg = (float)sums[C + c], db = (float)sums[c];
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
}

// ---------------------------------------------------------------------------------------------- affine + relu
// y = relu?(x * scale[c] + shift[c]) * (mask[pixel])   -> split store
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, int x_cs,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ shift, int relu,
                                                         const float* __restrict__ mask, SplitOut y, int y_cs,
                                                         long long npix, int C) {
    const int q = C >> 2;
    const long long total = npix * q;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 4;  // independent 16-byte loads in flight per thread
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
        float4 v[U];
        long long p[U];
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            p[u] = i / q;
            c[u] = (int)(i - p[u] * q) * 4;
            if (i < total) v[u] = *reinterpret_cast<const float4*>(x + p[u] * x_cs + c[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * stride >= total) break;
            float4 t = v[u];
            if (scale != nullptr) {
                const float4 s = *reinterpret_cast<const float4*>(scale + c[u]);
                t.x *= s.x; t.y *= s.y; t.z *= s.z; t.w *= s.w;
            }
            if (shift != nullptr) {
                const float4 b = *reinterpret_cast<const float4*>(shift + c[u]);
                t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
            }
            if (relu) {
                t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
            }
            if (mask != nullptr) {
                const float m = mask[p[u]];
                t.x *= m; t.y *= m; t.z *= m; t.w *= m;
            }
            store_split4(y, p[u] * y_cs + c[u], t);
        }
    }
}

// ---------------------------------------------------------------------------------------------- BN (train) apply
// bn_finalize fused into the apply pass: every block derives (scale, shift) of all C channels from the batch sums into
// shared memory (C <= 1024: a few hundred double ops per block), block 0 also publishes scale / shift / mean / invstd
// for the backward pass and performs the `n_updates` running-statistics updates. Then y = relu?(z*scale + shift).
__global__ void __launch_bounds__(256) bn_train_act_kernel(const float* __restrict__ z, int z_cs,
                                                           const double* __restrict__ sums, double count,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float momentum,
                                                           int n_updates, float* __restrict__ running_mean,
                                                           float* __restrict__ running_var, float* __restrict__ scale_out,
                                                           float* __restrict__ shift_out, float* __restrict__ mean_out,
                                                           float* __restrict__ invstd_out, int relu, SplitOut y, int y_cs,
                                                           long long npix, int C) {
    // every thread owns one channel quad (cq) and walks pixels with a constant stride: no index division in the loop,
    // the four (scale, shift) pairs live in registers
    const int q = C >> 2;
    const int cq = threadIdx.x % q, prow = threadIdx.x / q, ppb = blockDim.x / q;
    if (prow >= ppb) return;
    float sc[4], sh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = cq * 4 + e;
        const double m = sums[c] / count;
        double var = sums[C + c] / count - m * m;
        if (var < 0) var = 0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        sc[e] = g * invstd;
        sh[e] = b - (float)m * g * invstd;
        if (blockIdx.x == 0 && prow == 0) {
            scale_out[c] = sc[e];
            shift_out[c] = sh[e];
            if (mean_out) mean_out[c] = (float)m;
            if (invstd_out) invstd_out[c] = invstd;
            if (running_mean != nullptr && n_updates > 0) {
                const float unbiased = (float)(count > 1 ? var * count / (count - 1) : var);
                float rm = running_mean[c], rv = running_var[c];
                for (int i = 0; i < n_updates; ++i) {
                    rm = (1.f - momentum) * rm + momentum * (float)m;
                    rv = (1.f - momentum) * rv + momentum * unbiased;
                }
                running_mean[c] = rm;
                running_var[c] = rv;
            }
        }
    }
    constexpr int U = 8;
    const long long pstride = (long long)gridDim.x * ppb;
    for (long long p0 = (long long)blockIdx.x * ppb + prow; p0 < npix; p0 += pstride * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * pstride;
            if (p < npix) v[u] = *reinterpret_cast<const float4*>(z + p * z_cs + cq * 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * pstride;
            if (p >= npix) break;
            float4 t = v[u];
            t.x = t.x * sc[0] + sh[0]; t.y = t.y * sc[1] + sh[1]; t.z = t.z * sc[2] + sh[2]; t.w = t.w * sc[3] + sh[3];
            if (relu) {
                t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
            }
            store_split4(y, p * y_cs + cq * 4, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------- BN+ReLU backward
// g = dy * (z*scale+shift > 0); dz = gamma*invstd * (g - sum_g/m - zhat * sum_gz/m)        (train mode)
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const float* __restrict__ dy, int dy_cs,
                                                                const float* __restrict__ z, int z_cs,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd,
                                                                const double* __restrict__ sums, double count,
                                                                SplitOut dz, int dz_cs, long long npix, int C,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                int accumulate) {
    const int q = C >> 2;
    const int cq = threadIdx.x % q, prow = threadIdx.x / q, ppb = blockDim.x / q;
    if (prow >= ppb) return;
    float sc[4], sh[4], mu[4], is[4], mg[4], mgz[4];  // per-channel coefficients of my quad, in registers
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = cq * 4 + e;
        sc[e] = scale[c]; sh[e] = shift[c]; mu[e] = mean[c]; is[e] = invstd[c];
        const double sg = sums[c], sgz = sums[C + c];  // copy 0 holds the folded totals
        mg[e] = (float)(sg / count);
        mgz[e] = (float)(sgz / count);
        if (blockIdx.x == 0 && prow == 0) {  // dgamma = sum g*zhat, dbeta = sum g  (bn_bwd_finalize fused)
            const float dg = (float)sgz, db = (float)sg;
            if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
            if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
        }
    }
    constexpr int U = 4;
    const long long pstride = (long long)gridDim.x * ppb;
    for (long long p0 = (long long)blockIdx.x * ppb + prow; p0 < npix; p0 += pstride * U) {
        float4 d4[U], z4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * pstride;
            if (p < npix) {
                d4[u] = *reinterpret_cast<const float4*>(dy + p * dy_cs + cq * 4);
                z4[u] = *reinterpret_cast<const float4*>(z + p * z_cs + cq * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long p = p0 + u * pstride;
            if (p >= npix) break;
            const float d[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
            const float zz[4] = {z4[u].x, z4[u].y, z4[u].z, z4[u].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float g = (zz[e] * sc[e] + sh[e] > 0.f) ? d[e] : 0.f;
                const float zh = (zz[e] - mu[e]) * is[e];
                o[e] = sc[e] * (g - mg[e] - zh * mgz[e]);  // scale = gamma * invstd
            }
            store_split4(dz, p * dz_cs + cq * 4, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// g = dy * (y > 0) * mask[pixel]  -> split store        (bias+ReLU convs; mask multiply backward)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, int dy_cs,
                                                       const float* __restrict__ y, int y_cs,
                                                       const float* __restrict__ mask, SplitOut g, int g_cs,
                                                       long long npix, int C) {
    const int q = C >> 2;
    const long long total = npix * q;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / q;
        const int c = (int)(i - p * q) * 4;
        float4 d = *reinterpret_cast<const float4*>(dy + p * dy_cs + c);
        if (y != nullptr) {
            const float4 v = *reinterpret_cast<const float4*>(y + p * y_cs + c);
            d.x = v.x > 0.f ? d.x : 0.f;
            d.y = v.y > 0.f ? d.y : 0.f;
            d.z = v.z > 0.f ? d.z : 0.f;
            d.w = v.w > 0.f ? d.w : 0.f;
        }
        if (mask != nullptr) {
            const float m = mask[p];
            d.x *= m; d.y *= m; d.z *= m; d.w *= m;
        }
        store_split4(g, p * g_cs + c, d);
    }
}

// ---------------------------------------------------------------------------------------------- count_nonzero
__global__ void count_nonzero_kernel(const float* __restrict__ x, long long n4, unsigned long long* __restrict__ out) {
    unsigned int cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        cnt += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

static int ew_grid(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

// column-owner kernels: blockDim = q * (256 / q) threads (q = C/4 channel quads), each block iteration covers
// 256 / q pixels; grid sized so that every thread sees ~16 pixels, capped at 8 blocks per SM
static void col_launch_dims(long long npix, int C, int* threads, int* blocks) {
    const int q = C / 4;
    const int ppb = 256 / q > 0 ? 256 / q : 1;
    *threads = q * ppb;
    long long b = (npix + (long long)ppb * 16 - 1) / ((long long)ppb * 16);
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    *blocks = (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_split(const float* x, long long n, const a2x_output* out, a2x_stream_t stream) {
    A2X_REQUIRE(x && out && out->hi && out->b16 && n % 4 == 0, "split: bad args (n must be a multiple of 4)");
    split_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(x, n / 4, to_split(out));
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int check_c(int C) {
    if (C <= 0 || C % 4 != 0 || C > 1024) {
        set_error("channel count %d must be a multiple of 4 in (0, 1024]", C);
        return 1;
    }
    return 0;
}

int a2x_channel_stats(const float* x, int x_cs, long long npix, int C, double* sums, a2x_stream_t stream) {
    if (int r = check_c(C)) return r;
    A2X_REQUIRE(x && sums && npix > 0, "channel_stats: bad args");
    const int ppb = 256 / (C / 4);
    long long blocks = (npix + ppb - 1) / ppb;
    if (blocks > 148 * 2) blocks = 148 * 2;  // few blocks: the tail is 2C same-address double atomics per block
    channel_reduce_kernel<0><<<(int)blocks, 256, 2 * 256 * 4 * sizeof(float), (cudaStream_t)stream>>>(
        x, x_cs, nullptr, 0, nullptr, nullptr, nullptr, nullptr, npix, C, sums, 1);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps, float momentum,
                    int n_updates, float* running_mean, float* running_var, int C, float* scale, float* shift,
                    float* mean_out, float* invstd_out, a2x_stream_t stream) {
    A2X_REQUIRE(sums && scale && shift && C > 0 && count > 0, "bn_finalize: bad args");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, count, gamma, beta, eps, momentum,
                                                                        n_updates, running_mean, running_var, C, scale,
                                                                        shift, mean_out, invstd_out);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int C, float* scale, float* shift, a2x_stream_t stream) {
    A2X_REQUIRE(gamma && beta && running_mean && running_var && scale && shift && C > 0, "bn_eval_affine: bad args");
    bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var,
                                                                           eps, C, scale, shift);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_affine_act(const float* x, int x_cs, const float* scale, const float* shift, int relu, const float* mask,
                   const a2x_output* y, long long npix, int C, a2x_stream_t stream) {
    if (int r = check_c(C)) return r;
    A2X_REQUIRE(x && y && (y->hi || y->b16) && npix > 0, "affine_act: bad args");
    affine_act_kernel<<<ew_grid(npix * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, x_cs, scale, shift, relu, mask,
                                                                               to_split(y), y->cs, npix, C);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_bn_relu_bwd_reduce(const float* dy, int dy_cs, const float* z, int z_cs, const float* scale, const float* shift,
                           const float* mean, const float* invstd, long long npix, int C, double* sums,
                           a2x_stream_t stream) {
    if (int r = check_c(C)) return r;
    A2X_REQUIRE(dy && z && scale && shift && mean && invstd && sums && npix > 0, "bn_relu_bwd_reduce: bad args");
    const int ppb = 256 / (C / 4);
    long long blocks = (npix + (long long)ppb * 8 - 1) / ((long long)ppb * 8);
    if (blocks > 148 * 6) blocks = 148 * 6;  // the 2C double atomics per block land on A2X_BN_BWD_REPLICAS copies of `sums`
    if (blocks < 1) blocks = 1;
    channel_reduce_kernel<1><<<(int)blocks, 256, 2 * 256 * 4 * sizeof(float), (cudaStream_t)stream>>>(
        dy, dy_cs, z, z_cs, scale, shift, mean, invstd, npix, C, sums, A2X_BN_BWD_REPLICAS);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_bn_relu_bwd_apply(const float* dy, int dy_cs, const float* z, int z_cs, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const double* sums, double count,
                          const a2x_output* dz, long long npix, int C, float* dgamma, float* dbeta,
                          int accumulate_param_grads, a2x_stream_t stream) {
    if (int r = check_c(C)) return r;
    A2X_REQUIRE(dy && z && scale && shift && mean && invstd && sums && dz && (dz->hi || dz->b16) && npix > 0,
                "bn_relu_bwd_apply: bad args");
    int threads, blocks;
    col_launch_dims(npix, C, &threads, &blocks);
    bn_relu_bwd_apply_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        dy, dy_cs, z, z_cs, scale, shift, mean, invstd, sums, count, to_split(dz), dz->cs, npix, C, dgamma, dbeta,
        accumulate_param_grads);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_bn_train_act(const float* z, int z_cs, const double* sums, double count, const float* gamma, const float* beta,
                     float eps, float momentum, int n_updates, float* running_mean, float* running_var, float* scale,
                     float* shift, float* mean_out, float* invstd_out, int relu, const a2x_output* y, long long npix,
                     int C, a2x_stream_t stream) {
    if (int r = check_c(C)) return r;
    A2X_REQUIRE(z && sums && scale && shift && y && (y->hi || y->b16) && npix > 0 && count > 0, "bn_train_act: bad args");
    int threads, blocks;
    col_launch_dims(npix, C, &threads, &blocks);
    bn_train_act_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        z, z_cs, sums, count, gamma, beta, eps, momentum, n_updates, running_mean, running_var, scale, shift, mean_out,
        invstd_out, relu, to_split(y), y->cs, npix, C);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_relu_bwd(const float* dy, int dy_cs, const float* y, int y_cs, const float* mask, const a2x_output* g,
                 long long npix, int C, a2x_stream_t stream) {
    if (int r = check_c(C)) return r;
    A2X_REQUIRE(dy && g && g->hi && npix > 0, "relu_bwd: bad args");
    relu_bwd_kernel<<<ew_grid(npix * (C / 4)), 256, 0, (cudaStream_t)stream>>>(dy, dy_cs, y, y_cs, mask, to_split(g),
                                                                             g->cs, npix, C);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_sums_to_float(const double* sums, int C, float* out, int accumulate, a2x_stream_t stream) {
    A2X_REQUIRE(sums && out && C > 0, "sums_to_float: bad args");
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, C, nullptr, out, accumulate);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_count_nonzero(const float* x, long long n, unsigned long long* out, a2x_stream_t stream) {
    A2X_REQUIRE(x && out && n % 4 == 0, "count_nonzero: bad args");
    A2X_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    count_nonzero_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(x, n / 4, out);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
