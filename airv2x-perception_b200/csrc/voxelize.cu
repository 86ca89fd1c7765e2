// GPU voxelisation that reproduces the SEQUENTIAL first-come semantics of spconv's Point2VoxelCPU3d bit-exactly:
//   voxel id  = rank of the pillar's first point among all first points (input order), capped at max_voxels,
//   kept pts  = the <= 32 lowest-index points of the pillar, in input order, zero padded.
//
// Reference call sites: opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:59-72, :96-116 (third-party
// spconv arithmetic: c = floor((p - lo) / vs) in fp32, reject outside [0, grid)); collate at :142-175.
//
// Parallel formulation (all agents of a scene in one launch sequence, no host sync):
//   1. key kernel:   cell key per point, atomicMin(first[cell], i), atomicAdd(cnt[cell], 1)
//   2. rank kernel:  one CTA per agent: exclusive scan of "is first point" -> voxel id; per-voxel point counts ->
//                    exclusive scan -> CSR offsets; writes coords, per-agent voxel count
//   3. fill kernel:  every surviving point drops its index into its voxel's CSR segment (unordered)
//   4. gather kernel: one warp per voxel selects the 32 smallest indices with a bitonic network (order restored),
//                    gathers the points and writes the zero-padded [32][4] slab + num_points
// Output is a fixed-capacity slab per agent: voxels [n_agents][cap][32][4], coords [n_agents][cap][4]
// (agent,z,y,x), num_points [n_agents][cap], counts [n_agents] — consumed directly by the PFN kernels.
#include <limits.h>

#include "../../include/airv2x_b200.h"
#include "a2x_host.h"

namespace a2x {

extern int g_debug[16];

struct VoxGeom {
    float lo[3], hi[3], vs[3];
    int grid[3];  // nx, ny, nz
};

// project_points_by_matrix_torch (utils/box_utils.py:1038-1066): [x y z 1] . T^T in fp32. torch's CPU sgemm evaluates the
// K = 4 dot product as x*T0, then fused multiply-adds in k order (pinned bit-exact in tests/test_gpu_voxelize.py).
__device__ __forceinline__ float4 vox_project(float4 p, const float* __restrict__ T) {
    float4 o = p;
    o.x = __fadd_rn(__fmaf_rn(p.z, T[2], __fmaf_rn(p.y, T[1], __fmul_rn(p.x, T[0]))), T[3]);
    o.y = __fadd_rn(__fmaf_rn(p.z, T[6], __fmaf_rn(p.y, T[5], __fmul_rn(p.x, T[4]))), T[7]);
    o.z = __fadd_rn(__fmaf_rn(p.z, T[10], __fmaf_rn(p.y, T[9], __fmul_rn(p.x, T[8]))), T[11]);
    return o;
}

__global__ void __launch_bounds__(256) vox_key_kernel(const float* __restrict__ pts, const int* __restrict__ offsets,
                                                      const float* __restrict__ xf, VoxGeom g, int cells,
                                                      const unsigned char* __restrict__ ego_flags, int strict_range,
                                                      int* __restrict__ keys, int* __restrict__ first,
                                                      int* __restrict__ cnt) {
    const int a = blockIdx.y;
    const int p0 = offsets[a], np = offsets[a + 1] - p0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const float4 p = reinterpret_cast<const float4*>(pts)[p0 + i];   // sensor frame: the ego-box test below uses it
        const float4 pe = xf != nullptr ? vox_project(p, xf + a * 16) : p;
        const float v[3] = {pe.x, pe.y, pe.z};
        int c[3];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float q = floorf(__fdiv_rn(__fsub_rn(v[j], g.lo[j]), g.vs[j]));
            if (!(q >= 0.0f) || !(q < (float)g.grid[j])) ok = false;
            c[j] = (int)q;
        }
        // a1 point filters of the dataset (utils/pcd_utils.py:136-190): strictly inside the range; ego-box removal
        if (strict_range) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (!(v[j] > g.lo[j]) || !(v[j] < g.hi[j])) ok = false;
        }
        if (ego_flags != nullptr && ego_flags[a]) {
            if (p.x >= -1.95f && p.x <= 2.95f && p.y >= -1.1f && p.y <= 1.1f) ok = false;
        }
        int key = -1;
        if (ok) {
            key = (c[2] * g.grid[1] + c[1]) * g.grid[0] + c[0];
            atomicMin(&first[(long long)a * cells + key], i);
            atomicAdd(&cnt[(long long)a * cells + key], 1);
        }
        keys[p0 + i] = key;
    }
}

// block-wide exclusive scan of one int per thread (1024 threads)
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // protect warp_sums reuse across calls
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const int base = wid > 0 ? warp_sums[wid - 1] : 0;
    *total = warp_sums[31];
    return base + inc - v;
}

__global__ void __launch_bounds__(1024) vox_rank_kernel(const int* __restrict__ offsets, const int* __restrict__ keys,
                                                        const int* __restrict__ first, const int* __restrict__ cnt,
                                                        VoxGeom g, int cells, int cap, int max_voxels,
                                                        int* __restrict__ cell_vid, int* __restrict__ coords,
                                                        int* __restrict__ pcount, int* __restrict__ poff,
                                                        int* __restrict__ counts) {
    const int a = blockIdx.x;
    const int p0 = offsets[a], np = offsets[a + 1] - p0;
    const int chunk = (np + blockDim.x - 1) / blockDim.x;
    const int b = threadIdx.x * chunk, e = min(np, b + chunk);
    const int* fa = first + (long long)a * cells;
    const int* ca = cnt + (long long)a * cells;
    int* cv = cell_vid + (long long)a * cells;
    // pass 1: number of first points in my chunk
    int mine = 0;
    for (int i = b; i < e; ++i) {
        const int k = keys[p0 + i];
        if (k >= 0 && fa[k] == i) ++mine;
    }
    int total;
    int vid = block_excl_scan(mine, &total);
    const int nvox = min(total, max_voxels);
    for (int i = b; i < e; ++i) {
        const int k = keys[p0 + i];
        if (k >= 0 && fa[k] == i) {
            if (vid < max_voxels) {
                cv[k] = vid;
                const int cx = k % g.grid[0];
                const int cy = (k / g.grid[0]) % g.grid[1];
                const int cz = k / (g.grid[0] * g.grid[1]);
                int4 c4 = make_int4(a, cz, cy, cx);
                *reinterpret_cast<int4*>(coords + ((long long)a * cap + vid) * 4) = c4;
                pcount[(long long)a * cap + vid] = ca[k];
            } else {
                cv[k] = -1;  // pillar beyond the max_voxels cap: all of its points are dropped
            }
            ++vid;
        }
    }
    if (threadIdx.x == 0) counts[a] = nvox;
    __syncthreads();
    // CSR offsets over this agent's voxels
    const int vchunk = (nvox + blockDim.x - 1) / blockDim.x;
    const int vb = threadIdx.x * vchunk, ve = min(nvox, vb + vchunk);
    int s = 0;
    for (int v = vb; v < ve; ++v) s += pcount[(long long)a * cap + v];
    int tot2;
    int off = block_excl_scan(s, &tot2);
    for (int v = vb; v < ve; ++v) {
        poff[(long long)a * cap + v] = off;
        off += pcount[(long long)a * cap + v];
    }
}

// ---- parallel form of the ranking (the single-CTA kernel above took 0.21 ms of the step on 5 SMs):
// a two-level exclusive scan, VR_BLOCKS CTAs per agent, every thread owning a CONTIGUOUS chunk of the input order.
//   MODE 0: value(i) = [point i is the first point of its cell]   -> voxel ids in first-come order
//   MODE 1: value(v) = pcount[v]                                    -> CSR offsets of the voxels' point lists
constexpr int VR_BLOCKS = 64;
constexpr int VR_THREADS = 256;

struct VrArgs {
    const int* offsets; const int* keys; const int* first; const int* cnt;
    VoxGeom g;
    int cells, cap, max_voxels;
    int* cell_vid; int* coords; int* pcount; int* poff; int* counts;
    int* bsum;   // [agents][VR_BLOCKS] block totals, overwritten with exclusive block offsets by the scan kernel
};

template <int MODE>
__device__ __forceinline__ void vr_range(const VrArgs& p, int a, int& n, int& b, int& e) {
    n = MODE == 0 ? (p.offsets[a + 1] - p.offsets[a]) : p.counts[a];
    const int per_block = (n + VR_BLOCKS - 1) / VR_BLOCKS;
    const int chunk = (per_block + VR_THREADS - 1) / VR_THREADS;
    const int b0 = blockIdx.x * per_block;
    b = min(n, b0 + (int)threadIdx.x * chunk);
    e = min(min(n, b0 + per_block), b + chunk);
}

template <int MODE>
__device__ __forceinline__ int vr_local_sum(const VrArgs& p, int a, int b, int e) {
    int s = 0;
    if (MODE == 0) {
        const int p0 = p.offsets[a];
        const int* fa = p.first + (long long)a * p.cells;
        for (int i = b; i < e; ++i) {
            const int k = p.keys[p0 + i];
            s += (k >= 0 && fa[k] == i) ? 1 : 0;
        }
    } else {
        for (int v = b; v < e; ++v) s += p.pcount[(long long)a * p.cap + v];
    }
    return s;
}

__device__ __forceinline__ int block_excl_scan256(int v, int* total) {
    __shared__ int ws[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < wid) base += ws[w];
        tot += ws[w];
    }
    *total = tot;
    return base + inc - v;
}

template <int MODE>
__global__ void __launch_bounds__(VR_THREADS) vr_count_kernel(const VrArgs p) {
    const int a = blockIdx.y;
    int n, b, e;
    vr_range<MODE>(p, a, n, b, e);
    int total;
    block_excl_scan256(vr_local_sum<MODE>(p, a, b, e), &total);
    if (threadIdx.x == 0) p.bsum[a * VR_BLOCKS + blockIdx.x] = total;
}

// one warp-pair per agent: exclusive scan of the VR_BLOCKS block totals (in place); MODE 0 also publishes the voxel count
template <int MODE>
__global__ void __launch_bounds__(VR_BLOCKS) vr_scan_kernel(const VrArgs p) {
    __shared__ int s[VR_BLOCKS];
    const int a = blockIdx.x, t = threadIdx.x;
    s[t] = p.bsum[a * VR_BLOCKS + t];
    __syncthreads();
    int off = 0, tot = 0;
    for (int i = 0; i < VR_BLOCKS; ++i) {
        if (i < t) off += s[i];
        tot += s[i];
    }
    p.bsum[a * VR_BLOCKS + t] = off;
    if (MODE == 0 && t == 0) p.counts[a] = min(tot, p.max_voxels);
}

template <int MODE>
__global__ void __launch_bounds__(VR_THREADS) vr_assign_kernel(const VrArgs p) {
    const int a = blockIdx.y;
    int n, b, e;
    vr_range<MODE>(p, a, n, b, e);
    int total;
    int idx = block_excl_scan256(vr_local_sum<MODE>(p, a, b, e), &total) + p.bsum[a * VR_BLOCKS + blockIdx.x];
    if (MODE == 0) {
        const int p0 = p.offsets[a];
        const int* fa = p.first + (long long)a * p.cells;
        const int* ca = p.cnt + (long long)a * p.cells;
        int* cv = p.cell_vid + (long long)a * p.cells;
        for (int i = b; i < e; ++i) {
            const int k = p.keys[p0 + i];
            if (k >= 0 && fa[k] == i) {
                if (idx < p.max_voxels) {
                    cv[k] = idx;
                    const int cx = k % p.g.grid[0];
                    const int cy = (k / p.g.grid[0]) % p.g.grid[1];
                    const int cz = k / (p.g.grid[0] * p.g.grid[1]);
                    *reinterpret_cast<int4*>(p.coords + ((long long)a * p.cap + idx) * 4) = make_int4(a, cz, cy, cx);
                    p.pcount[(long long)a * p.cap + idx] = ca[k];
                } else {
                    cv[k] = -1;  // pillar beyond the max_voxels cap: all of its points are dropped
                }
                ++idx;
            }
        }
    } else {
        for (int v = b; v < e; ++v) {
            p.poff[(long long)a * p.cap + v] = idx;
            idx += p.pcount[(long long)a * p.cap + v];
        }
    }
}

__global__ void __launch_bounds__(256) vox_fill_kernel(const int* __restrict__ offsets, const int* __restrict__ keys,
                                                       const int* __restrict__ cell_vid, int cells, int cap,
                                                       const int* __restrict__ poff, int* __restrict__ fill,
                                                       int* __restrict__ plist) {
    const int a = blockIdx.y;
    const int p0 = offsets[a], np = offsets[a + 1] - p0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
        const int k = keys[p0 + i];
        if (k < 0) continue;
        const int v = cell_vid[(long long)a * cells + k];
        if (v < 0) continue;
        const int pos = atomicAdd(&fill[(long long)a * cap + v], 1);
        plist[p0 + poff[(long long)a * cap + v] + pos] = i;
    }
}

__device__ __forceinline__ int bitonic_sort32(int v, int lane) {  // ascending across lanes
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = ((lane & k) == 0);
            const bool lower = ((lane & j) == 0);
            v = (lower == up) ? min(v, o) : max(v, o);
        }
    }
    return v;
}
__device__ __forceinline__ int bitonic_merge32(int v, int lane) {  // input bitonic -> ascending
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const int o = __shfl_xor_sync(0xffffffffu, v, j);
        v = ((lane & j) == 0) ? min(v, o) : max(v, o);
    }
    return v;
}

__global__ void __launch_bounds__(256) vox_gather_kernel(const float* __restrict__ pts, const int* __restrict__ offsets,
                                                         const float* __restrict__ xf,
                                                         const int* __restrict__ counts, int cap,
                                                         const int* __restrict__ pcount, const int* __restrict__ poff,
                                                         const int* __restrict__ plist, float* __restrict__ voxels,
                                                         int* __restrict__ num_points) {
    const int a = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int nvox = counts[a];
    const int p0 = offsets[a];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < nvox; v += warps) {
        const long long vi = (long long)a * cap + v;
        const int n = pcount[vi];
        const int* lst = plist + p0 + poff[vi];
        int best = lane < n ? lst[lane] : INT_MAX;
        best = bitonic_sort32(best, lane);
        for (int base = 32; base < n; base += 32) {
            int c = base + lane < n ? lst[base + lane] : INT_MAX;
            c = bitonic_sort32(c, lane);
            const int rev = __shfl_sync(0xffffffffu, c, 31 - lane);  // descending
            best = min(best, rev);                                    // 32 smallest of the 64, bitonic
            best = bitonic_merge32(best, lane);
        }
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (best != INT_MAX) {
            p = reinterpret_cast<const float4*>(pts)[p0 + best];
            if (xf != nullptr) p = vox_project(p, xf + a * 16);
        }
        reinterpret_cast<float4*>(voxels)[vi * 32 + lane] = p;
        if (lane == 0) num_points[vi] = min(n, 32);
    }
}

}  // namespace a2x

using namespace a2x;

extern "C" {

size_t a2x_voxelize_workspace_bytes(int n_agents, long long total_points, int nx, int ny, int nz, int cap) {
    const size_t cells = (size_t)nx * ny * nz;
    size_t b = 0;
    b += (size_t)total_points * 4 * 2;      // keys, plist
    b += (size_t)n_agents * cells * 4 * 3;  // first, cnt, cell_vid
    b += (size_t)n_agents * cap * 4 * 3;    // pcount, poff, fill
    b += (size_t)n_agents * 64 * 4;         // block totals of the two-level ranking scan
    return b + 1024;
}

int a2x_voxelize(const float* points, const int* offsets_dev, int n_agents, long long total_points, const float* range6,
                 const float* vsize3, int max_points, int max_voxels, int cap, const unsigned char* ego_flags,
                 int strict_range, void* workspace, size_t workspace_bytes, float* voxels, int* coords, int* num_points,
                 int* counts, a2x_stream_t stream) {
    return a2x_voxelize_ex(points, offsets_dev, nullptr, n_agents, total_points, range6, vsize3, max_points, max_voxels, cap,
                           ego_flags, strict_range, workspace, workspace_bytes, voxels, coords, num_points, counts, stream);
}

int a2x_voxelize_ex(const float* points, const int* offsets_dev, const float* transforms_dev, int n_agents,
                    long long total_points, const float* range6, const float* vsize3, int max_points, int max_voxels,
                    int cap, const unsigned char* ego_flags, int strict_range, void* workspace, size_t workspace_bytes,
                    float* voxels, int* coords, int* num_points, int* counts, a2x_stream_t stream) {
    A2X_REQUIRE(points && offsets_dev && range6 && vsize3 && workspace && voxels && coords && num_points && counts,
                "voxelize: null argument");
    A2X_REQUIRE(max_points == 32, "voxelize: max_points_per_voxel must be 32 (one warp per pillar)");
    A2X_REQUIRE(n_agents > 0 && cap >= max_voxels && max_voxels > 0, "voxelize: bad sizes");
    VoxGeom g;
    for (int j = 0; j < 3; ++j) {
        g.lo[j] = range6[j];
        g.hi[j] = range6[3 + j];
        g.vs[j] = vsize3[j];
        g.grid[j] = (int)roundf((range6[3 + j] - range6[j]) / vsize3[j]);
    }
    const size_t cells = (size_t)g.grid[0] * g.grid[1] * g.grid[2];
    A2X_REQUIRE(cells < (size_t)INT_MAX, "voxelize: grid too large");
    A2X_REQUIRE(workspace_bytes >= a2x_voxelize_workspace_bytes(n_agents, total_points, g.grid[0], g.grid[1], g.grid[2], cap),
                "voxelize: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int* w = (int*)workspace;
    int* keys = w;
    int* plist = keys + total_points;
    int* first = plist + total_points;
    int* cnt = first + (size_t)n_agents * cells;
    int* cell_vid = cnt + (size_t)n_agents * cells;
    int* pcount = cell_vid + (size_t)n_agents * cells;
    int* poff = pcount + (size_t)n_agents * cap;
    int* fill = poff + (size_t)n_agents * cap;
    A2X_CHECK_CUDA(cudaMemsetAsync(first, 0x7f, (size_t)n_agents * cells * 4, st));
    A2X_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n_agents * cells * 4, st));
    A2X_CHECK_CUDA(cudaMemsetAsync(fill, 0, (size_t)n_agents * cap * 4, st));
    dim3 gk(64, n_agents);
    vox_key_kernel<<<gk, 256, 0, st>>>(points, offsets_dev, transforms_dev, g, (int)cells, ego_flags, strict_range, keys, first, cnt);
    A2X_LAUNCHED();
    if (g_debug[11] == 1) {  // reference single-CTA ranking (kept for A/B checks)
        vox_rank_kernel<<<n_agents, 1024, 0, st>>>(offsets_dev, keys, first, cnt, g, (int)cells, cap, max_voxels, cell_vid,
                                                  coords, pcount, poff, counts);
        A2X_LAUNCHED();
    } else {
        VrArgs va{offsets_dev, keys, first, cnt, g, (int)cells, cap, max_voxels, cell_vid, coords, pcount, poff, counts,
                  fill + (size_t)n_agents * cap};
        dim3 gv(VR_BLOCKS, n_agents);
        vr_count_kernel<0><<<gv, VR_THREADS, 0, st>>>(va);
        vr_scan_kernel<0><<<n_agents, VR_BLOCKS, 0, st>>>(va);
        vr_assign_kernel<0><<<gv, VR_THREADS, 0, st>>>(va);
        vr_count_kernel<1><<<gv, VR_THREADS, 0, st>>>(va);
        vr_scan_kernel<1><<<n_agents, VR_BLOCKS, 0, st>>>(va);
        vr_assign_kernel<1><<<gv, VR_THREADS, 0, st>>>(va);
        for (int i = 0; i < 6; ++i) A2X_LAUNCHED();
    }
    vox_fill_kernel<<<gk, 256, 0, st>>>(offsets_dev, keys, cell_vid, (int)cells, cap, poff, fill, plist);
    A2X_LAUNCHED();
    dim3 gg(128, n_agents);
    vox_gather_kernel<<<gg, 256, 0, st>>>(points, offsets_dev, transforms_dev, counts, cap, pcount, poff, plist, voxels, num_points);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
