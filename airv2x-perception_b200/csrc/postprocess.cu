// Detection decode + filters + rotated NMS on the GPU: replaces the host-side, shapely-based
// VoxelPostprocessor.post_process_airv2x of the reference for batch-1 inference.
//
// Reference semantics (order of operations):
//   data_utils/post_processor/voxel_postprocessor.py:666-840   post_process_airv2x
//   data_utils/post_processor/voxel_postprocessor.py:585-635   delta_to_boxes3d
//   utils/box_utils.py:195-258 (boxes_to_corners_3d, order "hwl"), utils/common_utils.py:60-82 (rotate_points_along_z)
//   utils/box_utils.py:981-1035 (remove_large_pred_bbx, remove_bbx_abnormal_z), :399-430 (range mask)
//   utils/box_utils.py:823-868 (nms_rotated: top-1000 by score, greedy, IoU of the first four corners' polygons)
//   utils/common_utils.py:150-191 (shapely polygon IoU) -> convex-quad clipping in double precision here.
//
// Pipeline (all stream-ordered, no host sync):
//   decode_kernel      one thread per anchor: objectness gate, box decode, 8 corners, size / z filters,
//                      warp-ballot compaction of the survivors into candidate slots (one atomic per warp)
//   select_sort_kernel one CTA: top-K (K = 1000) selection by score through a 2048-bin histogram, then a bitonic sort
//                      of the selected (score desc, anchor index asc) keys in shared memory
//   iou_mask_kernel    upper-triangular suppression bit matrix: IoU(i, j) > thr for the sorted candidates
//   sweep_kernel       one warp: greedy sweep over the bit matrix, range mask, ordered output
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"

#include <cuda_runtime.h>
#include <math.h>

namespace a2x {

constexpr int PP_TOP = 1000;       // nms_rotated: top = 1000
constexpr int PP_SORT_CAP = 4096;  // sort capacity (top-K plus the ties of the cut bin)
constexpr int PP_WORDS = (PP_TOP + 63) / 64;

struct PpCand {          // one decoded candidate
    float box[7];        // x, y, z, h, w, l, yaw  (order "hwl")
    float corner[24];    // 8 x (x, y, z)
    float score;
    int label;
    int anchor;          // flat anchor index (h, w, a)
    int in_range;        // all 8 corners inside the xy range (applied AFTER the NMS, like the reference)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ heads, int cs, int H, int W, int A,
                                                     int K, const float* __restrict__ anchors, float obj_thr,
                                                     float x_lo, float y_lo, float z_lo, float x_hi, float y_hi,
                                                     float z_hi, PpCand* __restrict__ cand, int cap,
                                                     int* __restrict__ count) {
    const int total = H * W * A;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false;
    PpCand c;
    if (i < total) {
        const int a = i % A;
        const int pix = i / A;
        const float* hp = heads + (long long)pix * cs;
        const int nc = A * K, nr = 7 * A;
        const float objectness = sigmoidf_(hp[nc + nr + a]);
        if (objectness > obj_thr) {
            const float* an = anchors + (long long)i * 7;
            const float* d = hp + nc + a * 7;
            const float diag = sqrtf(an[4] * an[4] + an[5] * an[5]);
            c.box[0] = d[0] * diag + an[0];
            c.box[1] = d[1] * diag + an[1];
            c.box[2] = d[2] * an[3] + an[2];
            c.box[3] = expf(d[3]) * an[3];
            c.box[4] = expf(d[4]) * an[4];
            c.box[5] = expf(d[5]) * an[5];
            c.box[6] = d[6] + an[6];
            // class label: psm viewed class-major (B, C, A, H, W); argmax over the non-background classes
            int best = 1;
            float bp = -1.f;
            for (int k = 1; k < K; ++k) {
                const float p = sigmoidf_(hp[k * A + a]);
                if (p > bp) {
                    bp = p;
                    best = k;
                }
            }
            c.label = best;
            c.score = objectness;
            c.anchor = i;
            // corners: (l, w, h) = (box5, box4, box3); points @ [[c, s, 0], [-s, c, 0], [0, 0, 1]] + centre
            const float l = c.box[5], w = c.box[4], h = c.box[3];
            const float cs_ = cosf(c.box[6]), sn = sinf(c.box[6]);
            const float tx[8] = {1, 1, -1, -1, 1, 1, -1, -1}, ty[8] = {-1, 1, 1, -1, -1, 1, 1, -1},
                        tz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
            float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float px = l * (tx[k] / 2), py = w * (ty[k] / 2), pz = h * (tz[k] / 2);
                const float rx = px * cs_ + py * (-sn), ry = px * sn + py * cs_;
                const float X = rx + c.box[0], Y = ry + c.box[1], Z = pz + c.box[2];
                c.corner[3 * k] = X; c.corner[3 * k + 1] = Y; c.corner[3 * k + 2] = Z;
                xmin = fminf(xmin, X); xmax = fmaxf(xmax, X);
                ymin = fminf(ymin, Y); ymax = fmaxf(ymax, Y);
                zmin = fminf(zmin, Z); zmax = fmaxf(zmax, Z);
            }
            const bool small = (xmax - xmin <= 6.f) && (ymax - ymin <= 6.f) && ((zmax - zmin) != 0.f);
            const bool zok = (zmin >= z_lo) && (zmax <= z_hi);
            c.in_range = (xmin >= x_lo && ymin >= y_lo && xmax <= x_hi && ymax <= y_hi) ? 1 : 0;
            keep = small && zok;
        }
    }
    // warp-ballot compaction: one atomic per warp, lanes write to consecutive slots
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m != 0) {
        int base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) {
            const int slot = base + __popc(m & ((1u << lane) - 1));
            if (slot < cap) cand[slot] = c;
        }
    }
}

// key order: score descending, anchor index ascending (deterministic tie-break)
__device__ __forceinline__ bool key_before(float sa, int ia, float sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(1024) select_sort_kernel(const PpCand* __restrict__ cand, const int* __restrict__ count,
                                                           int cap, float obj_thr, int* __restrict__ order,
                                                           int* __restrict__ n_sorted, int* __restrict__ status) {
    __shared__ int hist[2048];
    __shared__ float s_score[PP_SORT_CAP];
    __shared__ int s_idx[PP_SORT_CAP];
    __shared__ int s_cut, s_n;
    const int n = min(*count, cap);
    if (threadIdx.x == 0 && *count > cap) atomicOr(status, 1);  // candidate buffer overflow (reported, not hidden)
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) {
        s_cut = 0;
        s_n = 0;
    }
    __syncthreads();
    const float inv = 2047.f / (1.f - obj_thr);
    if (n > PP_TOP) {  // find the histogram bin holding the PP_TOP-th best score
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            int b = (int)((cand[i].score - obj_thr) * inv);
            b = max(0, min(2047, b));
            atomicAdd(&hist[b], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = 2047;
            for (; b > 0; --b) {
                acc += hist[b];
                if (acc >= PP_TOP) break;
            }
            s_cut = b;
        }
        __syncthreads();
    }
    const int cut = s_cut;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int b = (int)((cand[i].score - obj_thr) * inv);
        b = max(0, min(2047, b));
        if (b >= cut) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < PP_SORT_CAP) {
                s_score[slot] = cand[i].score;
                s_idx[slot] = i;
            }
        }
    }
    __syncthreads();
    int m = s_n;
    if (m > PP_SORT_CAP) {
        if (threadIdx.x == 0) atomicOr(status, 2);  // too many ties around the cut
        m = PP_SORT_CAP;
    }
    int p2 = 1;
    while (p2 < m) p2 <<= 1;
    for (int i = m + threadIdx.x; i < p2; i += blockDim.x) {
        s_score[i] = -INFINITY;
        s_idx[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const float sa = s_score[i], sb = s_score[ixj];
                    const int ia = s_idx[i] == 0x7fffffff ? 0x7fffffff : cand[s_idx[i]].anchor;
                    const int ib = s_idx[ixj] == 0x7fffffff ? 0x7fffffff : cand[s_idx[ixj]].anchor;
                    const bool a_first = key_before(sa, ia, sb, ib);
                    if (up != a_first) {
                        s_score[i] = sb; s_score[ixj] = sa;
                        const int t = s_idx[i]; s_idx[i] = s_idx[ixj]; s_idx[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    const int out_n = min(m, PP_TOP);
    for (int i = threadIdx.x; i < out_n; i += blockDim.x) order[i] = s_idx[i];
    if (threadIdx.x == 0) *n_sorted = out_n;
}

// ---- convex quadrilateral intersection (Sutherland-Hodgman) in double precision
__device__ double quad_area(const double* p, int n) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        a += p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
    }
    return 0.5 * a;
}

__device__ double quad_iou(const float* ca, const float* cb) {
    double A[8], B[8];
    for (int k = 0; k < 4; ++k) {
        A[2 * k] = ca[3 * k]; A[2 * k + 1] = ca[3 * k + 1];
        B[2 * k] = cb[3 * k]; B[2 * k + 1] = cb[3 * k + 1];
    }
    double sa = quad_area(A, 4), sb = quad_area(B, 4);
    if (sb < 0) {  // make the clip polygon counter-clockwise
        for (int k = 0; k < 2; ++k) {
            const double tx = B[2 * k], ty = B[2 * k + 1];
            B[2 * k] = B[2 * (3 - k)]; B[2 * k + 1] = B[2 * (3 - k) + 1];
            B[2 * (3 - k)] = tx; B[2 * (3 - k) + 1] = ty;
        }
        sb = -sb;
    }
    sa = fabs(sa);
    double poly[2][32];
    int n = 4, cur = 0;
    for (int k = 0; k < 8; ++k) poly[0][k] = A[k];
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double x1 = B[2 * e], y1 = B[2 * e + 1], x2 = B[2 * ((e + 1) & 3)], y2 = B[2 * ((e + 1) & 3) + 1];
        const double ex = x2 - x1, ey = y2 - y1;
        double* out = poly[cur ^ 1];
        const double* in = poly[cur];
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1) % n;
            const double px = in[2 * i], py = in[2 * i + 1], qx = in[2 * j], qy = in[2 * j + 1];
            const double dp = ex * (py - y1) - ey * (px - x1);  // >= 0: inside (left of the edge)
            const double dq = ex * (qy - y1) - ey * (qx - x1);
            if (dp >= 0) {
                out[2 * m] = px; out[2 * m + 1] = py; ++m;
            }
            if ((dp >= 0) != (dq >= 0)) {
                const double t = dp / (dp - dq);
                out[2 * m] = px + t * (qx - px); out[2 * m + 1] = py + t * (qy - py); ++m;
            }
        }
        n = m;
        cur ^= 1;
    }
    const double inter = n >= 3 ? fabs(quad_area(poly[cur], n)) : 0.0;
    const double uni = sa + sb - inter;
    return uni > 0 ? inter / uni : 0.0;
}

// mask[i][w] bit b set <=> j = 64 w + b > i and IoU(sorted i, sorted j) > thr
__global__ void __launch_bounds__(64) iou_mask_kernel(const PpCand* __restrict__ cand, const int* __restrict__ order,
                                                      const int* __restrict__ n_sorted, float thr,
                                                      unsigned long long* __restrict__ mask) {
    const int n = *n_sorted;
    const int i = blockIdx.x, w = blockIdx.y, b = threadIdx.x;
    const int j = w * 64 + b;
    bool sup = false;
    if (i < n && j < n && j > i) sup = (float)quad_iou(cand[order[i]].corner, cand[order[j]].corner) > thr;
    const unsigned lo = __ballot_sync(0xffffffffu, sup);
    __shared__ unsigned parts[2];
    if ((threadIdx.x & 31) == 0) parts[threadIdx.x >> 5] = lo;
    __syncthreads();
    if (threadIdx.x == 0 && i < n) mask[(long long)i * PP_WORDS + w] = (unsigned long long)parts[0] | ((unsigned long long)parts[1] << 32);
}

__global__ void __launch_bounds__(32) sweep_kernel(const PpCand* __restrict__ cand, const int* __restrict__ order,
                                                   const int* __restrict__ n_sorted,
                                                   const unsigned long long* __restrict__ mask, float* out_corners,
                                                   float* out_scores, int* out_labels, float* out_boxes,
                                                   int* out_anchor, int max_out, int* __restrict__ n_out) {
    __shared__ unsigned long long removed[PP_WORDS];
    const int lane = threadIdx.x;
    const int n = *n_sorted;
    for (int w = lane; w < PP_WORDS; w += 32) removed[w] = 0ull;
    __syncwarp();
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const bool dead = (removed[i >> 6] >> (i & 63)) & 1ull;
        if (!dead) {
            for (int w = lane; w < PP_WORDS; w += 32) removed[w] |= mask[(long long)i * PP_WORDS + w];
            const PpCand* c = cand + order[i];
            if (c->in_range) {
                if (kept < max_out) {
                    if (lane < 24) out_corners[kept * 24 + lane] = c->corner[lane];
                    if (lane < 7) out_boxes[kept * 7 + lane] = c->box[lane];
                    if (lane == 0) {
                        out_scores[kept] = c->score;
                        out_labels[kept] = c->label;
                        out_anchor[kept] = c->anchor;
                    }
                }
                ++kept;
            }
        }
        __syncwarp();
    }
    if (lane == 0) *n_out = min(kept, max_out);
}

// IoU matrix of two sets of quads (first four corners of (n, 8, 3) boxes): AP matching (eval_utils_opv2v.py:41-95)
__global__ void iou_matrix_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                  float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na * nb) return;
    out[i] = (float)quad_iou(a + (long long)(i / nb) * 24, b + (long long)(i % nb) * 24);
}

}  // namespace a2x

using namespace a2x;

extern "C" {

size_t a2x_postprocess_workspace_bytes(int n_anchors) {
    size_t cap = (size_t)n_anchors;
    return 256 + cap * sizeof(PpCand) + PP_TOP * sizeof(int) + (size_t)PP_TOP * PP_WORDS * sizeof(unsigned long long) + 256;
}

int a2x_postprocess_det(const float* heads, int heads_cs, int H, int W, int A, int num_class, const float* anchors,
                        float obj_threshold, float nms_threshold, const float* lidar_range6, void* workspace,
                        size_t workspace_bytes, float* out_corners, float* out_scores, int* out_labels, float* out_boxes,
                        int* out_anchor_index, int max_out, int* n_out_dev, int* status_dev, a2x_stream_t stream) {
    A2X_REQUIRE(heads && anchors && lidar_range6 && workspace && out_corners && out_scores && out_labels && out_boxes &&
                    out_anchor_index && n_out_dev && status_dev && H > 0 && W > 0 && A > 0 && num_class > 1 && max_out > 0,
                "postprocess_det: bad args");
    const int total = H * W * A;
    A2X_REQUIRE(workspace_bytes >= a2x_postprocess_workspace_bytes(total), "postprocess_det: workspace too small");
    A2X_REQUIRE(heads_cs >= A * num_class + 8 * A, "postprocess_det: heads need cls | reg | obj channels");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int* count = (int*)ws;
    int* n_sorted = count + 1;
    PpCand* cand = (PpCand*)(ws + 256);
    int* order = (int*)(ws + 256 + (size_t)total * sizeof(PpCand));
    unsigned long long* mask = (unsigned long long*)((((uintptr_t)(order + PP_TOP)) + 255) & ~(uintptr_t)255);
    A2X_CHECK_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    A2X_CHECK_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), st));
    decode_kernel<<<(total + 255) / 256, 256, 0, st>>>(heads, heads_cs, H, W, A, num_class, anchors, obj_threshold,
                                                      lidar_range6[0], lidar_range6[1], lidar_range6[2], lidar_range6[3],
                                                      lidar_range6[4], lidar_range6[5], cand, total, count);
    A2X_LAUNCHED();
    select_sort_kernel<<<1, 1024, 0, st>>>(cand, count, total, obj_threshold, order, n_sorted, status_dev);
    A2X_LAUNCHED();
    iou_mask_kernel<<<dim3(PP_TOP, PP_WORDS), 64, 0, st>>>(cand, order, n_sorted, nms_threshold, mask);
    A2X_LAUNCHED();
    sweep_kernel<<<1, 32, 0, st>>>(cand, order, n_sorted, mask, out_corners, out_scores, out_labels, out_boxes,
                                   out_anchor_index, max_out, n_out_dev);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_rotated_iou_matrix(const float* boxes_a, int na, const float* boxes_b, int nb, float* out, a2x_stream_t stream) {
    A2X_REQUIRE(boxes_a && boxes_b && out && na > 0 && nb > 0, "rotated_iou_matrix: bad args");
    iou_matrix_kernel<<<(na * nb + 127) / 128, 128, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, out);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
