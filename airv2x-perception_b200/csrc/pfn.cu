// PillarVFE (feature build + PFN linear/BN/ReLU/max over the 32 slots) fused with PointPillarScatter:
// one warp per pillar, lane = point slot, the 64-channel result goes straight to its BEV cell (one coalesced
// 256-byte NHWC row) — the [M,32,10] / [M,32,64] intermediates of the reference never exist.
//
// Reference: opencood/models/common_modules/airv2x_pillar_vfe.py:105-160 (features: xyz+i, xyz - pillar mean,
// xyz - voxel centre; padded slots zeroed), :27-49 (PFNLayer: Linear(10->64, no bias), BatchNorm1d(eps 1e-3,
// mom 0.01) over all M*32 rows INCLUDING the zero padded ones, ReLU, max over all 32 slots),
// point_pillar_scatter.py:15-82 (canvas[:, z + y*nx + x] = pillar).
//
// Train-mode batch statistics use linearity: sum_rows (W f) = W (sum f) and sum_rows (W f)^2 = W^T (sum f f^T) W,
// so one pass accumulates the 10-vector S1 and the 10x10 moment S2 and a 64-thread kernel finishes the job.
// The same moments close the BatchNorm backward (see pfn_bwd_finalize_kernel).
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"

namespace a2x {

struct PfnGeom {
    float vx, vy, vz, x_off, y_off, z_off;
    int nx, ny;
};

// Optional segmented addressing for voxeliser slabs: flat index p in [0, nseg*cap) -> slab seg_ids[p / cap],
// pillar p % cap, valid while < seg_counts[slab]. Null counts => plain [0, M) addressing.
struct PfnSeg {
    const int* ids;
    const int* counts;
    int cap, nseg;
};
__device__ __forceinline__ bool seg_resolve(const PfnSeg& sg, long long p, long long& pil) {
    if (sg.counts == nullptr) {
        pil = p;
        return true;
    }
    const int seg = (int)(p / sg.cap), idx = (int)(p % sg.cap);
    const int slab = sg.ids ? sg.ids[seg] : seg;
    pil = (long long)slab * sg.cap + idx;
    return idx < sg.counts[slab];
}
__device__ __forceinline__ double seg_rows(const PfnSeg& sg, double rows) {
    if (sg.counts == nullptr) return rows;
    double r = 0;
    for (int i = 0; i < sg.nseg; ++i) r += 32.0 * sg.counts[sg.ids ? sg.ids[i] : i];
    return r;
}

constexpr int NF = 10;    // point features
constexpr int NC = 64;    // PFN channels
constexpr int NS2 = 55;   // upper triangle of the 10x10 moment

// lane = slot: build the 10 features of this pillar's slot (zeros for padded slots)
__device__ __forceinline__ void pillar_features(const float* __restrict__ voxels, const int* __restrict__ num_points,
                                                const int* __restrict__ coords, const PfnGeom& g, long long pil,
                                                int lane, float f[NF], int& num, int& agent, int& cy, int& cx) {
    num = num_points[pil];
    const int4 c4 = *reinterpret_cast<const int4*>(coords + pil * 4);  // (agent, z, y, x)
    agent = c4.x;
    const int cz = c4.y;
    cy = c4.z;
    cx = c4.w;
    // (loading only the lane < num slots was measured SLOWER: it puts the point load behind the num_points load)
    const float4 pt = *reinterpret_cast<const float4*>(voxels + (pil * 32 + lane) * 4);
    float sx = pt.x, sy = pt.y, sz = pt.z;  // padded slots hold zeros (spconv zero-fills)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const float fn = (float)num;
    const float mx = sx / fn, my = sy / fn, mz = sz / fn;
    const float m = lane < num ? 1.f : 0.f;
    f[0] = pt.x * m;
    f[1] = pt.y * m;
    f[2] = pt.z * m;
    f[3] = pt.w * m;
    f[4] = (pt.x - mx) * m;
    f[5] = (pt.y - my) * m;
    f[6] = (pt.z - mz) * m;
    f[7] = (pt.x - ((float)cx * g.vx + g.x_off)) * m;
    f[8] = (pt.y - ((float)cy * g.vy + g.y_off)) * m;
    f[9] = (pt.z - ((float)cz * g.vz + g.z_off)) * m;
}

// ------------------------------------------------------------------------------------------------ moments
__global__ void __launch_bounds__(256) pfn_moments_kernel(const float* __restrict__ voxels,
                                                          const int* __restrict__ num_points,
                                                          const int* __restrict__ coords, PfnGeom g, long long M,
                                                          PfnSeg sg, double* __restrict__ moments /* [10 + 55] */) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float acc[NF + NS2];
#pragma unroll
    for (int i = 0; i < NF + NS2; ++i) acc[i] = 0.f;
    for (long long pp = warp0; pp < M; pp += nwarps) {
        long long pil;
        if (!seg_resolve(sg, pp, pil)) continue;
        float f[NF];
        int num, agent, cy, cx;
        pillar_features(voxels, num_points, coords, g, pil, lane, f, num, agent, cy, cx);
#pragma unroll
        for (int i = 0; i < NF; ++i) acc[i] += f[i];
        int t = NF;
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int j = i; j < NF; ++j) acc[t++] += f[i] * f[j];
    }
    __shared__ double red[NF + NS2];
    if (threadIdx.x < NF + NS2) red[threadIdx.x] = 0.0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NF + NS2; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(&red[i], (double)v);
    }
    __syncthreads();
    if (threadIdx.x < NF + NS2) atomicAdd(&moments[threadIdx.x], red[threadIdx.x]);
}

__device__ __forceinline__ int s2_index(int i, int j) {  // i <= j, row-major upper triangle
    return i * NF - (i * (i - 1)) / 2 + (j - i);
}

// batch statistics of y_c = W_c . f over `rows` rows, from the moments
__global__ void pfn_stats_finalize_kernel(const double* __restrict__ moments, double rows_in, PfnSeg sg,
                                          const float* __restrict__ W,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                          float momentum, int n_updates, float* __restrict__ running_mean,
                                          float* __restrict__ running_var, float* __restrict__ scale,
                                          float* __restrict__ shift, float* __restrict__ mean_out,
                                          float* __restrict__ invstd_out) {
    const int c = threadIdx.x;
    if (c >= NC) return;
    const double rows = seg_rows(sg, rows_in);
    double w[NF];
    for (int k = 0; k < NF; ++k) w[k] = (double)W[c * NF + k];
    double m = 0;
    for (int k = 0; k < NF; ++k) m += w[k] * moments[k];
    m /= rows;
    double e2 = 0;
    for (int i = 0; i < NF; ++i)
        for (int j = 0; j < NF; ++j) {
            const int a = i < j ? i : j, b = i < j ? j : i;
            e2 += w[i] * w[j] * moments[NF + s2_index(a, b)];
        }
    e2 /= rows;
    double var = e2 - m * m;
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    scale[c] = gamma[c] * invstd;
    shift[c] = beta[c] - (float)m * gamma[c] * invstd;
    if (mean_out) mean_out[c] = (float)m;
    if (invstd_out) invstd_out[c] = invstd;
    if (running_mean != nullptr && n_updates > 0) {
        const float unbiased = (float)(rows > 1 ? var * rows / (rows - 1) : var);
        float rm = running_mean[c], rv = running_var[c];
        for (int i = 0; i < n_updates; ++i) {
            rm = (1.f - momentum) * rm + momentum * (float)m;
            rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        running_mean[c] = rm;
        running_var[c] = rv;
    }
}

// ------------------------------------------------------------------------------------------------ forward + scatter
// One warp per pillar. Lane = point slot while the 10 features are built (pillar mean by shuffles); then lane = channel
// pair (c, c + 32): the warp walks over the pillar's REAL points only — a synthetic 60k-point cloud has ~2 points per
// pillar, the other 30 slots are zero rows whose PFN output is the per-channel constant relu(shift_c) — broadcasting a
// point's features by shuffle and keeping the running max / arg-max (lowest slot among ties, like a max over the 32
// slots in slot order; the padded slots enter once, as slot `num`). ~130 instructions per pillar instead of ~1300 for
// the all-slots butterfly. The arithmetic per (slot, channel) is unchanged: products in k order, fma(y, scale, shift).
__global__ void __launch_bounds__(256) pfn_scatter_kernel(const float* __restrict__ voxels,
                                                          const int* __restrict__ num_points,
                                                          const int* __restrict__ coords, PfnGeom g, long long M,
                                                          PfnSeg sg, const float* __restrict__ W,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          const int* __restrict__ agent_map, SplitOut canvas,
                                                          float* __restrict__ pillar_out /* [M][64] or null */,
                                                          unsigned char* __restrict__ amax /* [M][64] or null */,
                                                          unsigned long long* __restrict__ nz_count /* or null */) {
    const int lane = threadIdx.x & 31;
    float w0[NF], w1[NF];
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        w0[k] = W[lane * NF + k];
        w1[k] = W[(lane + 32) * NF + k];
    }
    const float sc0 = scale[lane], sc1 = scale[lane + 32], sh0 = shift[lane], sh1 = shift[lane + 32];
    const float pad0 = sh0 > 0.f ? sh0 : 0.f, pad1 = sh1 > 0.f ? sh1 : 0.f;   // PFN output of an all-zero (padded) slot
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned int nz = 0;
    for (long long pp = warp0; pp < M; pp += nwarps) {
        long long pil;
        if (!seg_resolve(sg, pp, pil)) continue;
        float f[NF];
        int num, agent, cy, cx;
        pillar_features(voxels, num_points, coords, g, pil, lane, f, num, agent, cy, cx);
        float best0 = -1.f, best1 = -1.f;   // outputs are >= +0 after the ReLU: the first slot always wins over the init
        int arg0 = 0, arg1 = 0;
        const int np = num < 32 ? num : 32;
        for (int s = 0; s < np; ++s) {
            float fs[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) fs[k] = __shfl_sync(0xffffffffu, f[k], s);
            float y0 = fs[0] * w0[0], y1 = fs[0] * w1[0];
#pragma unroll
            for (int k = 1; k < NF; ++k) {
                y0 = fmaf(fs[k], w0[k], y0);
                y1 = fmaf(fs[k], w1[k], y1);
            }
            y0 = fmaf(y0, sc0, sh0);
            y1 = fmaf(y1, sc1, sh1);
            y0 = y0 > 0.f ? y0 : 0.f;
            y1 = y1 > 0.f ? y1 : 0.f;
            if (y0 > best0) {
                best0 = y0;
                arg0 = s;
            }
            if (y1 > best1) {
                best1 = y1;
                arg1 = s;
            }
        }
        if (np < 32) {
            if (pad0 > best0) {
                best0 = pad0;
                arg0 = np;
            }
            if (pad1 > best1) {
                best1 = pad1;
                arg1 = np;
            }
        }
        nz += (best0 != 0.f) + (best1 != 0.f);
        const long long cell = ((long long)agent_map[agent] * g.ny + cy) * g.nx + cx;
        if (canvas.hi != nullptr) {
            float* o = canvas.hi + cell * NC;
            o[lane] = best0;
            o[lane + 32] = best1;
        }
        if (canvas.b16 != nullptr) {
            __nv_bfloat16* hb = canvas.b16 + cell * NC;
            __nv_bfloat16* lb = hb + canvas.ps;
            split_bf16(best0, hb[lane], lb[lane]);
            split_bf16(best1, hb[lane + 32], lb[lane + 32]);
        }
        if (pillar_out != nullptr) {
            pillar_out[pil * NC + lane] = best0;
            pillar_out[pil * NC + lane + 32] = best1;
        }
        if (amax != nullptr) {
            amax[pil * NC + lane] = (unsigned char)arg0;
            amax[pil * NC + lane + 32] = (unsigned char)arg1;
        }
    }
    if (nz_count != nullptr) {   // count_nonzero of the canvas (airv2x_where2com.py:122): every cell is written at most once
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
        if (lane == 0 && nz) atomicAdd(nz_count, (unsigned long long)nz);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// acc[c][0] = sum g, acc[c][1] = sum g*xhat, acc[c][2..11] = sum g*f   (g = dpillar routed through max and ReLU)
__global__ void __launch_bounds__(256) pfn_bwd_kernel(const float* __restrict__ voxels,
                                                      const int* __restrict__ num_points,
                                                      const int* __restrict__ coords, PfnGeom g, long long M,
                                                      PfnSeg sg, const float* __restrict__ W,
                                                      const float* __restrict__ scale,
                                                      const float* __restrict__ shift, const float* __restrict__ mean,
                                                      const float* __restrict__ invstd,
                                                      const int* __restrict__ agent_map,
                                                      const float* __restrict__ dcanvas,
                                                      const unsigned char* __restrict__ amax,
                                                      double* __restrict__ acc /* [64][12] */) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float a[2][12];
    float w[2][NF], sc[2], sh[2], mu[2], is[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
#pragma unroll
        for (int k = 0; k < NF; ++k) w[h][k] = W[c * NF + k];
        sc[h] = scale[c];
        sh[h] = shift[c];
        mu[h] = mean[c];
        is[h] = invstd[c];
#pragma unroll
        for (int k = 0; k < 12; ++k) a[h][k] = 0.f;
    }
    for (long long pp = warp0; pp < M; pp += nwarps) {
        long long pil;
        if (!seg_resolve(sg, pp, pil)) continue;
        float f[NF];
        int num, agent, cy, cx;
        pillar_features(voxels, num_points, coords, g, pil, lane, f, num, agent, cy, cx);
        const long long cell = ((long long)agent_map[agent] * g.ny + cy) * g.nx + cx;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            const float dp = dcanvas[cell * NC + c];
            const int src = amax[pil * NC + c];
            float fs[NF];
#pragma unroll
            for (int k = 0; k < NF; ++k) fs[k] = __shfl_sync(0xffffffffu, f[k], src);
            float x = 0.f;
#pragma unroll
            for (int k = 0; k < NF; ++k) x = fmaf(fs[k], w[h][k], x);
            const float y = fmaf(x, sc[h], sh[h]);
            const float gg = y > 0.f ? dp : 0.f;
            a[h][0] += gg;
            a[h][1] += gg * (x - mu[h]) * is[h];
#pragma unroll
            for (int k = 0; k < NF; ++k) a[h][2 + k] += gg * fs[k];
        }
    }
    // block-level reduction first: 8 warps -> one smem copy -> 768 global double atomics per block
    __shared__ float sacc[NC * 12];
    for (int i = threadIdx.x; i < NC * 12; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 12; ++k) atomicAdd(&sacc[(lane + 32 * h) * 12 + k], a[h][k]);
    __syncthreads();
    for (int i = threadIdx.x; i < NC * 12; i += blockDim.x)
        if (sacc[i] != 0.f) atomicAdd(&acc[i], (double)sacc[i]);
}

// dW_c = gamma*invstd * (G - (A/m) S1 - (Bz/m) * invstd * (S2 W_c - mu S1)),  dgamma = Bz, dbeta = A
__global__ void pfn_bwd_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ moments,
                                        double rows_in, PfnSeg sg, const float* __restrict__ W, const float* __restrict__ scale,
                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                        float* __restrict__ dW, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                        int accumulate) {
    const int c = threadIdx.x;
    if (c >= NC) return;
    const double rows = seg_rows(sg, rows_in);
    const double A = acc[c * 12], Bz = acc[c * 12 + 1];
    const double is = (double)invstd[c], mu = (double)mean[c], gs = (double)scale[c];  // scale = gamma*invstd
    for (int k = 0; k < NF; ++k) {
        double s2w = 0;
        for (int j = 0; j < NF; ++j) {
            const int a = k < j ? k : j, b = k < j ? j : k;
            s2w += moments[NF + s2_index(a, b)] * (double)W[c * NF + j];
        }
        const double sum_xhat_f = is * (s2w - mu * moments[k]);
        const double v = gs * (acc[c * 12 + 2 + k] - (A / rows) * moments[k] - (Bz / rows) * sum_xhat_f);
        dW[c * NF + k] = accumulate ? dW[c * NF + k] + (float)v : (float)v;
    }
    dgamma[c] = accumulate ? dgamma[c] + (float)Bz : (float)Bz;
    dbeta[c] = accumulate ? dbeta[c] + (float)A : (float)A;
}

static int warp_grid(long long M, int cap_blocks = 148 * 8) {
    long long b = (M + 7) / 8;  // 8 warps per block
    if (b > cap_blocks) b = cap_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

static PfnSeg make_seg(const a2x_pfn_segments* s, long long* m) {
    PfnSeg r{nullptr, nullptr, 0, 0};
    if (s != nullptr && s->seg_counts != nullptr) {
        r.ids = s->seg_ids;
        r.counts = s->seg_counts;
        r.cap = s->seg_cap;
        r.nseg = s->nseg;
        *m = (long long)s->seg_cap * s->nseg;
    }
    return r;
}

static PfnGeom make_geom(const a2x_pfn_geom* g) {
    PfnGeom r;
    r.vx = g->voxel_x;
    r.vy = g->voxel_y;
    r.vz = g->voxel_z;
    r.x_off = g->x_offset;
    r.y_off = g->y_offset;
    r.z_off = g->z_offset;
    r.nx = g->nx;
    r.ny = g->ny;
    return r;
}

int a2x_pfn_moments(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                    const a2x_pfn_segments* seg, double* moments65, a2x_stream_t stream) {
    const PfnSeg sg = make_seg(seg, &m);
    A2X_REQUIRE(voxels && num_points && coords && geom && moments65 && m > 0, "pfn_moments: bad args");
    A2X_CHECK_CUDA(cudaMemsetAsync(moments65, 0, sizeof(double) * (NF + NS2), (cudaStream_t)stream));
    pfn_moments_kernel<<<warp_grid(m, 148 * 2), 256, 0, (cudaStream_t)stream>>>(voxels, num_points, coords, make_geom(geom), m,
                                                                     sg, moments65);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_pfn_stats_finalize(const double* moments65, double rows, const a2x_pfn_segments* seg, const float* w,
                           const float* gamma, const float* beta,
                           float eps, float momentum, int n_updates, float* running_mean, float* running_var,
                           float* scale, float* shift, float* mean_out, float* invstd_out, a2x_stream_t stream) {
    long long m_unused = 1;
    const PfnSeg sg = make_seg(seg, &m_unused);
    A2X_REQUIRE(moments65 && w && gamma && beta && scale && shift && (rows > 0 || sg.counts), "pfn_stats_finalize: bad args");
    pfn_stats_finalize_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(moments65, rows, sg, w, gamma, beta, eps, momentum,
                                                                n_updates, running_mean, running_var, scale, shift,
                                                                mean_out, invstd_out);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_pfn_scatter(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                    const a2x_pfn_segments* seg, const float* w, const float* scale, const float* shift,
                    const int* agent_map, const a2x_output* canvas, float* pillar_out, unsigned char* amax,
                    a2x_stream_t stream) {
    return a2x_pfn_scatter_ex(voxels, num_points, coords, m, geom, seg, w, scale, shift, agent_map, canvas, pillar_out, amax,
                              nullptr, stream);
}

int a2x_pfn_scatter_ex(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                       const a2x_pfn_segments* seg, const float* w, const float* scale, const float* shift,
                       const int* agent_map, const a2x_output* canvas, float* pillar_out, unsigned char* amax,
                       long long* nonzero_count, a2x_stream_t stream) {
    const PfnSeg sg = make_seg(seg, &m);
    A2X_REQUIRE(voxels && num_points && coords && geom && w && scale && shift && agent_map && canvas &&
                    (canvas->hi || canvas->b16) && m > 0,
                "pfn_scatter: bad args");
    SplitOut so;
    so.hi = canvas->hi;
    so.b16 = (__nv_bfloat16*)canvas->b16;
    so.ps = canvas->b16_plane;
    pfn_scatter_kernel<<<warp_grid(m), 256, 0, (cudaStream_t)stream>>>(voxels, num_points, coords, make_geom(geom), m,
                                                                     sg, w, scale, shift, agent_map, so, pillar_out, amax,
                                                                     (unsigned long long*)nonzero_count);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_pfn_bwd(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                const a2x_pfn_segments* seg, const float* w, const float* scale, const float* shift, const float* mean, const float* invstd,
                const int* agent_map, const float* dcanvas, const unsigned char* amax, const double* moments65,
                double rows, double* acc_ws /* 64*12 doubles */, float* dw, float* dgamma, float* dbeta, int accumulate,
                a2x_stream_t stream) {
    const PfnSeg sg = make_seg(seg, &m);
    A2X_REQUIRE(voxels && num_points && coords && geom && w && scale && shift && mean && invstd && agent_map && dcanvas &&
                    amax && moments65 && acc_ws && dw && dgamma && dbeta && m > 0,
                "pfn_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(acc_ws, 0, sizeof(double) * NC * 12, st));
    pfn_bwd_kernel<<<warp_grid(m, 148 * 4), 256, 0, st>>>(voxels, num_points, coords, make_geom(geom), m, sg, w, scale, shift,
                                                mean, invstd, agent_map, dcanvas, amax, acc_ws);
    A2X_LAUNCHED();
    pfn_bwd_finalize_kernel<<<1, 64, 0, st>>>(acc_ws, moments65, rows, sg, w, scale, mean, invstd, dw, dgamma, dbeta,
                                             accumulate);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
