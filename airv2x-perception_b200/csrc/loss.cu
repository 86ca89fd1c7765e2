// Fused detection loss (forward value + gradient w.r.t. the head logits) — one HBM pass over the head output.
//
// Reference: opencood/loss/point_pillar_loss_multiclass.py:96-179 (forward), :181-215 (sigmoid focal, alpha .25,
// gamma 2, anchor-major [B,H,W,A,K] view), :13-75 (weighted smooth-L1, beta 1/9), :273-289 (sin-difference on yaw),
// :161-166 (BCE objectness with the 1e-6 guards). total = reg*reg_coe + cls*cls_weight + obj.
//
// Head layout here: NHWC [B,H,W,heads_cs] with channels [0, A*K) = psm (a*K+k), [A*K, A*K+7A) = rm (a*7+j),
// [A*K+7A, A*K+8A) = obj — exactly the reference's permute(0,2,3,1) views of psm / rm / obj.
//
// LEGACY = true: PointPillarLoss of the legacy `point_pillar_*` models (opencood/loss/point_pillar_loss.py:77-215): one
// logit per anchor (K = 1, target = pos_equal_one), the focal term is divided by B once (the multi-class variant divides
// twice), the same smooth-L1, no objectness head ([0, A) = psm, [A, 8A) = rm).
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"

namespace a2x {

__global__ void count_pos_kernel(const float* __restrict__ pos, long long per_sample, int B, float* __restrict__ npos) {
    const int b = blockIdx.y;
    float c = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample;
         i += (long long)gridDim.x * blockDim.x)
        c += pos[b * per_sample + i] > 0.f ? 1.f : 0.f;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c != 0.f) atomicAdd(&npos[b], c);
}

// one thread per (b, h, w); A anchors, K classes
template <bool LEGACY>
__global__ void __launch_bounds__(256) det_loss_kernel(const float* __restrict__ heads, int cs, int A, int K,
                                                       const float* __restrict__ targets,   // [B,HW,A*7]
                                                       const float* __restrict__ pos,       // [B,HW,A]
                                                       const int* __restrict__ class_ids,   // [B,HW,A]
                                                       const float* __restrict__ npos, int B, long long HW,
                                                       float cls_weight, float reg_coe, float* __restrict__ dheads,
                                                       int dcs, double* __restrict__ loss3 /* reg, cls, obj */) {
    const long long total = (long long)B * HW;
    double l_reg = 0, l_cls = 0, l_obj = 0;
    const float obj_norm = 1.f / (float)(total * A);
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(p / HW);
        const float* hrow = heads + p * cs;
        float* drow = dheads ? dheads + p * dcs : nullptr;
        const float wn = 1.f / fmaxf(npos[b], 1.f);
        for (int a = 0; a < A; ++a) {
            const float pm = pos[p * A + a];
            const bool is_pos = pm > 0.f;
            const int cid = LEGACY ? (is_pos ? 0 : -1) : class_ids[p * A + a];
            // ---- classification: sigmoid focal loss, weight 1/max(npos,1) for every anchor
            for (int k = 0; k < K; ++k) {
                const float x = hrow[a * K + k];
                const float t = (k == cid) ? 1.f : 0.f;
                const float pr = 1.f / (1.f + expf(-x));
                const float aw = t * 0.25f + (1.f - t) * 0.75f;
                const float pt = t * (1.f - pr) + (1.f - t) * pr;
                const float fw = aw * pt * pt;
                const float bce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
                l_cls += (double)(fw * bce * wn);
                if (drow) {
                    const float dpt = (1.f - 2.f * t) * pr * (1.f - pr);
                    const float g = wn * (aw * 2.f * pt * dpt * bce + fw * (pr - t));
                    drow[a * K + k] = g * cls_weight / (LEGACY ? (float)B : (float)B * (float)B);
                }
            }
            // ---- regression: smooth-L1 with sin-difference on yaw, positives only
            const float rw = is_pos ? wn : 0.f;
            for (int j = 0; j < 7; ++j) {
                const float r = hrow[A * K + a * 7 + j];
                const float g = targets[(p * A + a) * 7 + j];
                float diff, ddiff;
                if (j == 6) {
                    float b1 = sinf(r) * cosf(g), b2 = cosf(r) * sinf(g);
                    if (isnan(b2)) b2 = b1;
                    diff = b1 - b2;
                    ddiff = cosf(r) * cosf(g) + sinf(r) * sinf(g);
                    if (isnan(g)) ddiff = 0.f;
                } else {
                    diff = isnan(g) ? 0.f : r - g;
                    ddiff = isnan(g) ? 0.f : 1.f;
                }
                const float n = fabsf(diff);
                const float beta = 1.0f / 9.0f;
                const float l = n < beta ? 0.5f * n * n / beta : n - 0.5f * beta;
                l_reg += (double)(l * rw);
                if (drow) {
                    const float dl = n < beta ? diff / beta : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
                    drow[A * K + a * 7 + j] = dl * ddiff * rw * reg_coe / (float)B;
                }
            }
            // ---- objectness BCE (mean over all B*H*W*A)
            if (!LEGACY) {
                const float o = hrow[A * K + A * 7 + a];
                const float s = 1.f / (1.f + expf(-o));
                const float l = -(pm * logf(s + 1e-6f) + (1.f - pm) * logf(1.f - s + 1e-6f));
                l_obj += (double)l;
                if (drow) {
                    const float ds = s * (1.f - s);
                    drow[A * K + A * 7 + a] = -(pm * ds / (s + 1e-6f) - (1.f - pm) * ds / (1.f - s + 1e-6f)) * obj_norm;
                }
            }
        }
        if (drow) {
            for (int c = A * K + A * (LEGACY ? 7 : 8); c < dcs; ++c) drow[c] = 0.f;
        }
    }
    // block reduce
    __shared__ double red[3][8];
    for (int o = 16; o > 0; o >>= 1) {
        l_reg += __shfl_xor_sync(0xffffffffu, l_reg, o);
        l_cls += __shfl_xor_sync(0xffffffffu, l_cls, o);
        l_obj += __shfl_xor_sync(0xffffffffu, l_obj, o);
    }
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][wid] = l_reg;
        red[1][wid] = l_cls;
        red[2][wid] = l_obj;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0, c = 0, o = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            r += red[0][i];
            c += red[1][i];
            o += red[2][i];
        }
        atomicAdd(&loss3[0], r * reg_coe / B);
        atomicAdd(&loss3[1], c * cls_weight / (LEGACY ? (double)B : (double)B * B));
        atomicAdd(&loss3[2], o * obj_norm);
    }
}

}  // namespace a2x

using namespace a2x;

extern "C" {

/* loss3[0..2] = (reg, cls, obj) terms (doubles, zeroed here); total = their sum. dheads may be NULL (value only).
 * npos_ws: B floats of workspace. */
static int det_loss_launch(bool legacy, const float* heads, int heads_cs, int B, long long HW, int A, int K, const float* targets,
                           const float* pos_equal_one, const int* class_ids, float cls_weight, float reg_coe, float* npos_ws,
                           float* dheads, int dheads_cs, double* loss3, a2x_stream_t stream);

int a2x_det_loss(const float* heads, int heads_cs, int B, long long HW, int A, int K, const float* targets,
                 const float* pos_equal_one, const int* class_ids, float cls_weight, float reg_coe, float* npos_ws,
                 float* dheads, int dheads_cs, double* loss3, a2x_stream_t stream) {
    A2X_REQUIRE(class_ids, "det_loss: bad args");
    return det_loss_launch(false, heads, heads_cs, B, HW, A, K, targets, pos_equal_one, class_ids, cls_weight, reg_coe, npos_ws,
                           dheads, dheads_cs, loss3, stream);
}

/* PointPillarLoss of the legacy models: loss3 = (reg, conf, 0) */
int a2x_det_loss_legacy(const float* heads, int heads_cs, int B, long long HW, int A, const float* targets,
                        const float* pos_equal_one, float cls_weight, float reg_coe, float* npos_ws, float* dheads,
                        int dheads_cs, double* loss3, a2x_stream_t stream) {
    return det_loss_launch(true, heads, heads_cs, B, HW, A, 1, targets, pos_equal_one, nullptr, cls_weight, reg_coe, npos_ws,
                           dheads, dheads_cs, loss3, stream);
}

static int det_loss_launch(bool legacy, const float* heads, int heads_cs, int B, long long HW, int A, int K, const float* targets,
                           const float* pos_equal_one, const int* class_ids, float cls_weight, float reg_coe, float* npos_ws,
                           float* dheads, int dheads_cs, double* loss3, a2x_stream_t stream) {
    A2X_REQUIRE(heads && targets && pos_equal_one && npos_ws && loss3 && B > 0 && HW > 0 && A > 0 && K > 0,
                "det_loss: bad args");
    const int need = A * K + (legacy ? 7 : 8) * A;
    A2X_REQUIRE(heads_cs >= need && (!dheads || dheads_cs >= need), "det_loss: head stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(npos_ws, 0, sizeof(float) * B, st));
    A2X_CHECK_CUDA(cudaMemsetAsync(loss3, 0, sizeof(double) * 3, st));
    dim3 g1(32, B);
    count_pos_kernel<<<g1, 256, 0, st>>>(pos_equal_one, HW * A, B, npos_ws);
    A2X_LAUNCHED();
    long long blocks = (B * HW + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (legacy)
        det_loss_kernel<true><<<(int)blocks, 256, 0, st>>>(heads, heads_cs, A, K, targets, pos_equal_one, class_ids, npos_ws, B,
                                                          HW, cls_weight, reg_coe, dheads, dheads_cs, loss3);
    else
        det_loss_kernel<false><<<(int)blocks, 256, 0, st>>>(heads, heads_cs, A, K, targets, pos_equal_one, class_ids, npos_ws, B,
                                                           HW, cls_weight, reg_coe, dheads, dheads_cs, loss3);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
