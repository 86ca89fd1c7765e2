// Anchor-target assignment on the GPU: replaces VoxelPostprocessor.generate_label_airv2x
// (opencood/data_utils/post_processor/voxel_postprocessor.py:217-354) + bbox_overlaps (opencood/utils/box_overlaps.pyx:17-56)
// for a whole batch. HBM-bound integer / compare work: N = H*W*A anchors x n ground-truth boxes per sample.
//
//   assign_scan_kernel   thread = anchor: IoU against every ground truth of the sample (staged in shared memory), keeps
//                        the first gt above the positive threshold and whether all IoUs are below the negative one;
//                        per gt, the (IoU, lowest anchor index) maximum over all anchors via a packed 64-bit atomicMax
//   assign_write_kernel  thread = anchor: match = first positive gt, else the lowest gt this anchor is the best anchor
//                        of (np.unique keeps the first occurrence, :292-294); writes pos / neg / class / 7 regression
//                        targets (double arithmetic, stored as the fp32 the loss kernel reads)
//
// IoU arithmetic is the Cython build's: "+ 1" is a double literal, so the sums / products around it run in double and
// are rounded when stored to the float variables (oracle/labels_oracle.py:bbox_overlaps, pinned to the real reference).
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"

namespace a2x {

__device__ __forceinline__ float label_iou(const float4 b, const float4 q, float area_q) {
    const float iw = (float)__dadd_rn((double)__fsub_rn(fminf(b.z, q.z), fmaxf(b.x, q.x)), 1.0);
    if (!(iw > 0.f)) return 0.f;
    const float ih = (float)__dadd_rn((double)__fsub_rn(fminf(b.w, q.w), fmaxf(b.y, q.y)), 1.0);
    if (!(ih > 0.f)) return 0.f;
    const float inter = __fmul_rn(iw, ih);
    const double area_b = __dmul_rn(__dadd_rn((double)__fsub_rn(b.z, b.x), 1.0), __dadd_rn((double)__fsub_rn(b.w, b.y), 1.0));
    const float ua = (float)__dsub_rn(__dadd_rn(area_b, (double)area_q), (double)inter);
    return __fdiv_rn(inter, ua);
}

constexpr int LBL_CHUNK = 128;

__global__ void __launch_bounds__(256) assign_scan_kernel(const float4* __restrict__ anchor_box, int N,
                                                          const float4* __restrict__ gt_box, const int* __restrict__ gt_off,
                                                          float pos_thr, float neg_thr, int* __restrict__ code,
                                                          unsigned long long* __restrict__ best) {
    __shared__ float4 sq[LBL_CHUNK];
    __shared__ float sa[LBL_CHUNK];
    __shared__ unsigned long long sbest[LBL_CHUNK];
    const int b = blockIdx.y;
    const int g0 = gt_off[b], n = gt_off[b + 1] - g0;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < N;
    const float4 box = live ? anchor_box[a] : make_float4(0.f, 0.f, 0.f, 0.f);
    int first_pos = -1;
    bool all_neg = true;
    for (int c0 = 0; c0 < n; c0 += LBL_CHUNK) {
        const int m = min(LBL_CHUNK, n - c0);
        __syncthreads();
        if (threadIdx.x < m) {
            const float4 q = gt_box[g0 + c0 + threadIdx.x];
            sq[threadIdx.x] = q;
            sa[threadIdx.x] = (float)__dmul_rn(__dadd_rn((double)__fsub_rn(q.z, q.x), 1.0), __dadd_rn((double)__fsub_rn(q.w, q.y), 1.0));
            sbest[threadIdx.x] = 0ull;
        }
        __syncthreads();
        if (live) {
            for (int k = 0; k < m; ++k) {
                const float iou = label_iou(box, sq[k], sa[k]);
                if (iou > pos_thr && first_pos < 0) first_pos = c0 + k;
                if (!(iou < neg_thr)) all_neg = false;
                if (iou > 0.f)
                    atomicMax(&sbest[k], ((unsigned long long)__float_as_uint(iou) << 32) | (0xffffffffu - (unsigned)a));
            }
        }
        __syncthreads();
        if (threadIdx.x < m && sbest[threadIdx.x] != 0ull) atomicMax(&best[g0 + c0 + threadIdx.x], sbest[threadIdx.x]);
    }
    if (live) code[(long long)b * N + a] = first_pos >= 0 ? first_pos : (all_neg ? -2 : -1);
}

__global__ void __launch_bounds__(256) assign_write_kernel(const double* __restrict__ anchors, int N,
                                                           const double* __restrict__ gt7, const int* __restrict__ gt_cls,
                                                           const int* __restrict__ gt_off, const int* __restrict__ code,
                                                           const unsigned long long* __restrict__ best,
                                                           float* __restrict__ targets, float* __restrict__ pos,
                                                           float* __restrict__ neg, int* __restrict__ cls) {
    const int b = blockIdx.y;
    const int g0 = gt_off[b], n = gt_off[b + 1] - g0;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    const long long o = (long long)b * N + a;
    const int c = code[o];
    int g = c >= 0 ? c : -1;
    bool highest = false;
    const unsigned long long* bb = best + g0;
    for (int k = 0; k < n; ++k) {
        const unsigned long long v = bb[k];
        if (v != 0ull && (0xffffffffu - (unsigned)(v & 0xffffffffull)) == (unsigned)a) {
            highest = true;
            if (g < 0) g = k;
        }
    }
    pos[o] = g >= 0 ? 1.f : 0.f;
    neg[o] = (c == -2 && !highest) ? 1.f : 0.f;
    cls[o] = g >= 0 ? gt_cls[g0 + g] : 0;
    float t[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g >= 0) {
        const double* an = anchors + (long long)a * 7;
        const double* gt = gt7 + (long long)(g0 + g) * 7;
        const double d = sqrt(__dadd_rn(__dmul_rn(an[4], an[4]), __dmul_rn(an[5], an[5])));
        t[0] = (float)((gt[0] - an[0]) / d);
        t[1] = (float)((gt[1] - an[1]) / d);
        t[2] = (float)((gt[2] - an[2]) / an[3]);
        t[3] = (float)log(gt[3] / an[3]);
        t[4] = (float)log(gt[4] / an[4]);
        t[5] = (float)log(gt[5] / an[5]);
        t[6] = (float)(gt[6] - an[6]);
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) targets[o * 7 + j] = t[j];
}

}  // namespace a2x

extern "C" int a2x_assign_targets(const float* anchor_standup, const double* anchors, int n_anchors, const float* gt_standup,
                                  const double* gt_boxes, const int* gt_class, const int* gt_offsets_dev, int total_gt,
                                  int B, float pos_threshold, float neg_threshold, int* code_ws,
                                  unsigned long long* best_ws, float* targets, float* pos_equal_one, float* neg_equal_one,
                                  int* class_ids, a2x_stream_t stream) {
    A2X_REQUIRE(anchor_standup && anchors && gt_offsets_dev && code_ws && best_ws && targets && pos_equal_one &&
                    neg_equal_one && class_ids && n_anchors > 0 && B > 0 && total_gt >= 0,
                "assign_targets: bad args");
    A2X_REQUIRE(total_gt == 0 || (gt_standup && gt_boxes && gt_class), "assign_targets: null ground-truth arrays");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(best_ws, 0, (size_t)(total_gt > 0 ? total_gt : 1) * sizeof(unsigned long long), st));
    dim3 grid((n_anchors + 255) / 256, B);
    a2x::assign_scan_kernel<<<grid, 256, 0, st>>>((const float4*)anchor_standup, n_anchors, (const float4*)gt_standup,
                                                 gt_offsets_dev, pos_threshold, neg_threshold, code_ws, best_ws);
    A2X_LAUNCHED();
    a2x::assign_write_kernel<<<grid, 256, 0, st>>>(anchors, n_anchors, gt_boxes, gt_class, gt_offsets_dev, code_ws, best_ws,
                                                  targets, pos_equal_one, neg_equal_one, class_ids);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}
