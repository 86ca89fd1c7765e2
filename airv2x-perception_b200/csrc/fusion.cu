// Where2comm communication mask and per-pixel cross-agent attention fusion (forward + backward). HBM-bound.
//
// Reference: opencood/models/where2comm_modules/where2comm_fuse.py
//   Communication.forward  :83-149  (sigmoid-max confidence, 5x5 Gaussian conv, threshold / top-K, rate, ego = 1)
//   AttentionFusion        :152-164 + ScaledDotProductAttention :14-45 (per pixel, ego row only)
#include <float.h>

#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"

namespace a2x {

// ---------------------------------------------------------------------------------------------- confidence map
// conf[n,h,w] = max_c sigmoid(psm[n,h,w,c]), c < ncls  (== sigmoid(max logit), sigmoid is monotone)
__global__ void conf_map_kernel(const float* __restrict__ psm, int cs, int ncls, long long npix,
                                float* __restrict__ conf) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const float* r = psm + p * cs;
        float m = -FLT_MAX;
        for (int c = 0; c < ncls; ++c) m = fmaxf(m, r[c]);
        conf[p] = 1.f / (1.f + expf(-m));
    }
}

// smoothed = conv2d(conf, w[k x k], zero pad) + bias ; eval mode: mask = smoothed > thr
__global__ void gauss_mask_kernel(const float* __restrict__ conf, const float* __restrict__ wk, const float* __restrict__ bias,
                                  int ksz, int N, int H, int W, float thr, int write_mask, float* __restrict__ smooth,
                                  float* __restrict__ mask) {
    const long long total = (long long)N * H * W;
    const int r = ksz / 2;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(p % W);
        const int h = (int)((p / W) % H);
        const long long n = p / ((long long)W * H);
        float acc = 0.f;
        if (ksz > 0) {
            for (int i = 0; i < ksz; ++i) {
                const int hh = h + i - r;
                if (hh < 0 || hh >= H) continue;
                for (int j = 0; j < ksz; ++j) {
                    const int ww = w + j - r;
                    if (ww < 0 || ww >= W) continue;
                    acc += wk[i * ksz + j] * conf[(n * H + hh) * W + ww];
                }
            }
            acc += bias[0];
        } else {
            acc = conf[p];
        }
        smooth[p] = acc;
        if (write_mask) mask[p] = acc > thr ? 1.f : 0.f;
    }
}

// Train mode: per agent, mask = 1 on the K largest smoothed values (ties -> lowest index), radix select on the
// order-preserving uint image of the floats. One CTA per agent.
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024) topk_mask_kernel(const float* __restrict__ smooth, int HW,
                                                         const int* __restrict__ k_per_agent, float* __restrict__ mask) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_remaining, s_tie_budget;
    const int n = blockIdx.x;
    const float* v = smooth + (long long)n * HW;
    float* m = mask + (long long)n * HW;
    const int K = k_per_agent[n];
    if (K <= 0) {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) m[i] = 0.f;
        return;
    }
    if (K >= HW) {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) m[i] = 1.f;
        return;
    }
    if (threadIdx.x == 0) {
        s_prefix = 0;
        s_remaining = (unsigned)K;
    }
    __syncthreads();
    // find the K-th largest key, 8 bits at a time from the top
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t pmask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = threadIdx.x; i < HW; i += blockDim.x) {
            const uint32_t key = f2ord(v[i]);
            if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int rem = s_remaining;
            int d = 255;
            for (; d >= 0; --d) {
                if (hist[d] >= rem) break;
                rem -= hist[d];
            }
            s_prefix = prefix | ((uint32_t)d << shift);
            s_remaining = rem;  // how many of the elements equal (so far) to the prefix are still needed
        }
        __syncthreads();
    }
    const uint32_t kth = s_prefix;
    if (threadIdx.x == 0) s_tie_budget = s_remaining;
    __syncthreads();
    // elements > kth are in; elements == kth: the first `s_remaining` by index (sequential scan by thread 0 is
    // avoided: ties are rare; resolve them with an ordered pass over chunks)
    for (int base = 0; base < HW; base += blockDim.x) {
        const int i = base + threadIdx.x;
        uint32_t key = 0;
        bool tie = false;
        if (i < HW) {
            key = f2ord(v[i]);
            tie = (key == kth);
            if (!tie) m[i] = key > kth ? 1.f : 0.f;
        }
        // ordered tie resolution within the chunk
        const unsigned ball = __ballot_sync(0xffffffffu, tie);
        __shared__ unsigned int warp_cnt[32];
        const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) warp_cnt[wid] = __popc(ball);
        __syncthreads();
        if (tie) {
            unsigned before = __popc(ball & ((1u << lane) - 1u));
            for (int w2 = 0; w2 < wid; ++w2) before += warp_cnt[w2];
            m[i] = before < s_tie_budget ? 1.f : 0.f;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) tot += warp_cnt[w2];
            s_tie_budget = tot >= s_tie_budget ? 0u : s_tie_budget - tot;
        }
        __syncthreads();
    }
}

// per-scene: ones[b] = sum(mask over the scene's agents) (before the ego override), then mask[ego agent] = 1
__global__ void mask_rate_ego_kernel(float* __restrict__ mask, int HW, const int* __restrict__ scene_start,
                                     const int* __restrict__ scene_len, float* __restrict__ ones) {
    const int b = blockIdx.y;
    const int s0 = scene_start[b], n = scene_len[b];
    const long long total = (long long)n * HW;
    float cnt = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float* mp = mask + (long long)s0 * HW + i;
        cnt += *mp;
        if (i < HW) *mp = 1.f;  // ego = first agent of the scene
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt != 0.f) atomicAdd(&ones[b], cnt);
}

// ---------------------------------------------------------------------------------------------- mask resize
// F.interpolate(mask, size, mode="bilinear", align_corners=False) of the communication mask when the level-0 features
// are finer than the confidence map (legacy stride-2 shrink header, where2comm_fuse.py:230-236)
__global__ void resize_bilinear_kernel(const float* __restrict__ src, int h, int w, float* __restrict__ dst, int H, int W,
                                       long long total) {
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(i % W);
        const int Y = (int)((i / W) % H);
        const long long n = i / ((long long)W * H);
        float fy = sh * (Y + 0.5f) - 0.5f, fx = sw * (X + 0.5f) - 0.5f;
        fy = fy < 0.f ? 0.f : fy;
        fx = fx < 0.f ? 0.f : fx;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
        const float* s = src + n * h * w;
        dst[i] = hy * (hx * s[y0 * w + x0] + lx * s[y0 * w + x1]) + ly * (hx * s[y1 * w + x0] + lx * s[y1 * w + x1]);
    }
}

// ---------------------------------------------------------------------------------------------- sparse feature select
// Where2comm transmits only the BEV cells its communication mask selected (where2comm_fuse.py:83-149, :237). Sender:
// warp-ballot compaction of the selected cells of one agent's level-0 map into (count, cell index, feature row)
// records. Receiver: the records of every agent are pulled through a pointer table (local memory after an all-gather,
// or PEER GPU memory over NVLink: system-scope loads) and scattered into the dense, zero-filled per-agent maps the
// fusion kernels read. Bytes on the wire = mask rate x dense size; no host round trip for the counts.
__global__ void __launch_bounds__(256) mask_compact_kernel(const float* __restrict__ x, int x_cs,
                                                           const float* __restrict__ mask, int force_all, int hw,
                                                           int C, int* __restrict__ hdr, int* __restrict__ idx,
                                                           float* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int q = C >> 2;
    for (int p0 = warp * 32; p0 < hw; p0 += nwarps * 32) {
        const int p = p0 + lane;
        const bool nz = p < hw && mask[p] != 0.f;
        const bool sel = p < hw && (force_all || nz);
        const unsigned ms = __ballot_sync(0xffffffffu, sel), mz = __ballot_sync(0xffffffffu, nz);
        int base = 0;
        if (lane == 0) {
            if (ms) base = atomicAdd(&hdr[0], __popc(ms));
            if (mz) atomicAdd(&hdr[1], __popc(mz));  // cells selected by the mask itself (the communication rate)
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (sel) idx[base + __popc(ms & ((1u << lane) - 1))] = p;
        unsigned rem = ms;
        int r = 0;
        while (rem) {  // the warp copies the selected rows one by one (C/4 float4 per row)
            const int b = __ffs(rem) - 1;
            rem &= rem - 1;
            const float4* src = reinterpret_cast<const float4*>(x + (long long)(p0 + b) * x_cs);
            float4* dst = reinterpret_cast<float4*>(vals + (long long)(base + r) * C);
            for (int c = lane; c < q; c += 32) dst[c] = src[c];
            ++r;
        }
    }
}

__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_sys_i32(const int* p) {
    int v;
    asm volatile("ld.global.relaxed.sys.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// dst[a][idx][:] = vals[a][r][:] for r < count[a]; dst is zero-filled by the caller. blockIdx.y = agent.
__global__ void __launch_bounds__(256) mask_decompact_ptrs_kernel(const char* const* __restrict__ bufs, long long off_idx,
                                                                  long long off_vals, int C, int hw,
                                                                  float* __restrict__ dst) {
    const int a = blockIdx.y;
    const char* base = bufs[a];
    const int count = min(ld_sys_i32(reinterpret_cast<const int*>(base)), hw);
    const int* idx = reinterpret_cast<const int*>(base + off_idx);
    const float* vals = reinterpret_cast<const float*>(base + off_vals);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int q = C >> 2;
    for (int r = warp; r < count; r += nwarps) {
        const int p = ld_sys_i32(idx + r);
        const float4* src = reinterpret_cast<const float4*>(vals + (long long)r * C);
        float4* out = reinterpret_cast<float4*>(dst + ((long long)a * hw + p) * C);
        for (int c = lane; c < q; c += 32) out[c] = ld_sys_f4(src + c);
    }
}

// ---------------------------------------------------------------------------------------------- attention fusion
// One thread group of G = min(32, C/4) lanes per pixel, V = C/(4G) float4 per lane.
template <int G, int V>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}

constexpr int ATT_MAX_AGENTS = 16;

template <int G, int V>
__global__ void __launch_bounds__(256) att_fuse_fwd_kernel(const float* __restrict__ x, int HW, int C, int n_agents,
                                                           float inv_sqrt_c, SplitOut out) {
    const int gid_raw = (blockIdx.x * blockDim.x + threadIdx.x) / G;  // pixel
    const int gl = threadIdx.x % G;
    const bool active = gid_raw < HW;  // inactive groups compute on a clamped pixel (full-warp shuffles stay legal)
    const int gid = active ? gid_raw : HW - 1;
    const float* x0 = x + (long long)gid * C;
    float4 q[V];
#pragma unroll
    for (int v = 0; v < V; ++v) q[v] = *reinterpret_cast<const float4*>(x0 + (v * G + gl) * 4);
    float s[ATT_MAX_AGENTS];
    float smax = -FLT_MAX;
    for (int j = 0; j < n_agents; ++j) {
        const float* xj = x + ((long long)j * HW + gid) * C;
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 k = *reinterpret_cast<const float4*>(xj + (v * G + gl) * 4);
            d += q[v].x * k.x + q[v].y * k.y + q[v].z * k.z + q[v].w * k.w;
        }
        d = group_sum<G, V>(d) * inv_sqrt_c;
        s[j] = d;
        smax = fmaxf(smax, d);
    }
    float den = 0.f;
    for (int j = 0; j < n_agents; ++j) {
        s[j] = expf(s[j] - smax);
        den += s[j];
    }
    const float inv = 1.f / den;
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0, 0, 0, 0);
    for (int j = 0; j < n_agents; ++j) {
        const float a = s[j] * inv;
        const float* xj = x + ((long long)j * HW + gid) * C;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 k = *reinterpret_cast<const float4*>(xj + (v * G + gl) * 4);
            acc[v].x += a * k.x; acc[v].y += a * k.y; acc[v].z += a * k.z; acc[v].w += a * k.w;
        }
    }
    if (!active) return;
#pragma unroll
    for (int v = 0; v < V; ++v) store_split4(out, (long long)gid * C + (v * G + gl) * 4, acc[v]);
}

// backward: given dout (per pixel, C), write dx for every agent of the scene.
//   a = softmax(s), s_j = <x0, xj>/sqrt(C), out = sum_j a_j xj
//   da_j = <dout, xj>; ds_j = a_j (da_j - sum_k a_k da_k)
//   dx_j = a_j dout + ds_j x0 / sqrt(C)  (j >= 0)   and   dx_0 += sum_j ds_j xj / sqrt(C)
template <int G, int V>
__global__ void __launch_bounds__(256) att_fuse_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                           int HW, int C, int n_agents, float inv_sqrt_c,
                                                           float* __restrict__ dx) {
    const int gid_raw = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const bool active = gid_raw < HW;
    const int gid = active ? gid_raw : HW - 1;
    const float* x0 = x + (long long)gid * C;
    float4 q[V], go[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        q[v] = *reinterpret_cast<const float4*>(x0 + (v * G + gl) * 4);
        go[v] = *reinterpret_cast<const float4*>(dout + (long long)gid * C + (v * G + gl) * 4);
    }
    float s[ATT_MAX_AGENTS], da[ATT_MAX_AGENTS];
    float smax = -FLT_MAX;
    for (int j = 0; j < n_agents; ++j) {
        const float* xj = x + ((long long)j * HW + gid) * C;
        float d = 0.f, e = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 k = *reinterpret_cast<const float4*>(xj + (v * G + gl) * 4);
            d += q[v].x * k.x + q[v].y * k.y + q[v].z * k.z + q[v].w * k.w;
            e += go[v].x * k.x + go[v].y * k.y + go[v].z * k.z + go[v].w * k.w;
        }
        s[j] = group_sum<G, V>(d) * inv_sqrt_c;
        da[j] = group_sum<G, V>(e);
        smax = fmaxf(smax, s[j]);
    }
    float den = 0.f;
    for (int j = 0; j < n_agents; ++j) {
        s[j] = expf(s[j] - smax);
        den += s[j];
    }
    const float inv = 1.f / den;
    float dot = 0.f;
    for (int j = 0; j < n_agents; ++j) {
        s[j] *= inv;
        dot += s[j] * da[j];
    }
    float4 d0[V];
#pragma unroll
    for (int v = 0; v < V; ++v) d0[v] = make_float4(0, 0, 0, 0);
    for (int j = 0; j < n_agents; ++j) {
        const float a = s[j];
        const float ds = a * (da[j] - dot) * inv_sqrt_c;
        const float* xj = x + ((long long)j * HW + gid) * C;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 k = *reinterpret_cast<const float4*>(xj + (v * G + gl) * 4);
            d0[v].x += ds * k.x; d0[v].y += ds * k.y; d0[v].z += ds * k.z; d0[v].w += ds * k.w;
            float4 r = make_float4(a * go[v].x + ds * q[v].x, a * go[v].y + ds * q[v].y, a * go[v].z + ds * q[v].z,
                                   a * go[v].w + ds * q[v].w);
            if (j > 0) {
                if (active) *reinterpret_cast<float4*>(dx + ((long long)j * HW + gid) * C + (v * G + gl) * 4) = r;
            } else {
                d0[v].x += r.x; d0[v].y += r.y; d0[v].z += r.z; d0[v].w += r.w;
            }
        }
    }
    if (!active) return;
#pragma unroll
    for (int v = 0; v < V; ++v) *reinterpret_cast<float4*>(dx + (long long)gid * C + (v * G + gl) * 4) = d0[v];
}

static int grid1d(long long total, int cap = 148 * 16) {
    long long b = (total + 255) / 256;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_comm_confidence(const float* psm, int psm_cs, int ncls, long long npix, float* conf, a2x_stream_t stream) {
    A2X_REQUIRE(psm && conf && ncls > 0 && npix > 0, "comm_confidence: bad args");
    conf_map_kernel<<<grid1d(npix), 256, 0, (cudaStream_t)stream>>>(psm, psm_cs, ncls, npix, conf);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_comm_smooth_mask(const float* conf, const float* gauss_w, const float* gauss_b, int ksz, int n, int h, int w,
                         float threshold, int write_mask, float* smooth, float* mask, a2x_stream_t stream) {
    A2X_REQUIRE(conf && smooth && (ksz == 0 || (gauss_w && gauss_b)) && (!write_mask || mask), "comm_smooth_mask: bad args");
    gauss_mask_kernel<<<grid1d((long long)n * h * w), 256, 0, (cudaStream_t)stream>>>(conf, gauss_w, gauss_b, ksz, n, h, w,
                                                                                   threshold, write_mask, smooth, mask);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_comm_topk_mask(const float* smooth, int n, int hw, const int* k_per_agent, float* mask, a2x_stream_t stream) {
    A2X_REQUIRE(smooth && k_per_agent && mask && n > 0 && hw > 0, "comm_topk_mask: bad args");
    topk_mask_kernel<<<n, 1024, 0, (cudaStream_t)stream>>>(smooth, hw, k_per_agent, mask);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_resize_bilinear(const float* src, int n, int h, int w, float* dst, int H, int W, a2x_stream_t stream) {
    A2X_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && H > 0 && W > 0, "resize_bilinear: bad args");
    const long long total = (long long)n * H * W;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    resize_bilinear_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, h, w, dst, H, W, total);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_mask_compact(const float* x, int x_cs, const float* mask, int force_all, int hw, int C, int* hdr, int* idx,
                     float* vals, a2x_stream_t stream) {
    A2X_REQUIRE(x && mask && hdr && idx && vals && hw > 0 && C > 0 && C % 4 == 0 && x_cs >= C, "mask_compact: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(hdr, 0, 2 * sizeof(int), st));
    int blocks = (hw + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    mask_compact_kernel<<<blocks, 256, 0, st>>>(x, x_cs, mask, force_all, hw, C, hdr, idx, vals);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_mask_decompact_ptrs(const void* const* bufs_dev, long long off_idx_bytes, long long off_vals_bytes, int n_agents,
                            int hw, int C, float* dst, a2x_stream_t stream) {
    A2X_REQUIRE(bufs_dev && dst && n_agents > 0 && hw > 0 && C > 0 && C % 4 == 0 && off_idx_bytes % 4 == 0 &&
                    off_vals_bytes % 16 == 0,
                "mask_decompact_ptrs: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(dst, 0, (size_t)n_agents * hw * C * sizeof(float), st));
    mask_decompact_ptrs_kernel<<<dim3(148, n_agents), 256, 0, st>>>((const char* const*)bufs_dev, off_idx_bytes,
                                                                   off_vals_bytes, C, hw, dst);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_comm_rate_ego(float* mask, int hw, int n_scenes, const int* scene_start, const int* scene_len, float* ones,
                      a2x_stream_t stream) {
    A2X_REQUIRE(mask && scene_start && scene_len && ones && n_scenes > 0, "comm_rate_ego: bad args");
    A2X_CHECK_CUDA(cudaMemsetAsync(ones, 0, sizeof(float) * n_scenes, (cudaStream_t)stream));
    dim3 grid(64, n_scenes);
    mask_rate_ego_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, hw, scene_start, scene_len, ones);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

#define A2X_ATT_DISPATCH(KERNEL, ...)                                                         \
    do {                                                                                      \
        const long long threads = (long long)hw * g;                                          \
        const int blocks = (int)((threads + 255) / 256);                                      \
        if (c == 64) KERNEL<16, 1><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);    \
        else if (c == 128) KERNEL<32, 1><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        else if (c == 256) KERNEL<32, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        else if (c == 32) KERNEL<8, 1><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);  \
        else if (c == 384) KERNEL<32, 3><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        else if (c == 512) KERNEL<32, 4><<<blocks, 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
        else {                                                                                \
            set_error("attention fusion: unsupported channel count %d", c);                   \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

/* x: [n_agents][hw][c] dense NHWC of ONE scene (agent 0 = ego); out: [hw][c] */
int a2x_att_fuse_fwd(const float* x, int n_agents, int hw, int c, const a2x_output* out, a2x_stream_t stream) {
    A2X_REQUIRE(x && out && out->hi && n_agents > 0 && n_agents <= ATT_MAX_AGENTS && hw > 0, "att_fuse_fwd: bad args (<= 16 agents)");
    const int g = c / 4 < 32 ? c / 4 : 32;
    const float isc = 1.0f / sqrtf((float)c);
    SplitOut so;
    so.hi = out->hi;
    so.b16 = (__nv_bfloat16*)out->b16;
    so.ps = out->b16_plane;
    A2X_ATT_DISPATCH(att_fuse_fwd_kernel, x, hw, c, n_agents, isc, so);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_att_fuse_bwd(const float* x, const float* dout, int n_agents, int hw, int c, float* dx, a2x_stream_t stream) {
    A2X_REQUIRE(x && dout && dx && n_agents > 0 && n_agents <= ATT_MAX_AGENTS && hw > 0, "att_fuse_bwd: bad args");
    const int g = c / 4 < 32 ? c / 4 : 32;
    const float isc = 1.0f / sqrtf((float)c);
    A2X_ATT_DISPATCH(att_fuse_bwd_kernel, x, dout, hw, c, n_agents, isc, dx);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
