/*
 * airv2x_b200.h — C ABI of the B200-native collaborative-perception hot path.
 *
 * Every entry point is a stream-ordered launcher over raw DEVICE pointers (plain pointers and sizes, no torch
 * types), returns 0 on success (1 = bad argument, 2 = CUDA error; message via a2x_last_error(), thread-local),
 * performs no hidden allocation (callers pass outputs and workspaces) and never synchronises the device.
 * Activations are fp32 NHWC ("pixel-major": [n][h][w][c], c contiguous) with an explicit pixel stride
 * (`*_cs`, elements) so that channel slices of a wider buffer (the 384-channel concat) can be read/written in
 * place. Dense contractions run as tcgen05.mma.kind::f16 on bf16 split operands (v = h + l; h*h + l*h + h*l, three
 * MMAs into one fp32 TMEM accumulator: fp32-equivalent to ~1e-4 on the logits); a2x_output / a2x_operand carry the two
 * bf16 planes next to the fp32 value.
 *
 * Each section cites the reference interface it replaces (paths relative to the reference repo root).
 */
#ifndef AIRV2X_B200_H
#define AIRV2X_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* a2x_stream_t; /* cudaStream_t */

/* ---------------------------------------------------------------- library */
const char* a2x_last_error(void);
int a2x_version(void);
/* Bring-up / A-B switches (all 0 = production behaviour; used by tests and profiling scripts only):
 *   1 tile width override of the tap-GEMM (N per CTA)      7 = 1: disable the halo (3x3 stride-1 A-tile reuse) variants
 *   8 cap on persistent CTAs (default 148 = one per SM)    10 K-split override of the weight-gradient kernel
 *   11 = 1: single-CTA reference ranking in the voxeliser   2, 3, 5, 6, 9: UMMA descriptor / epilogue experiments */
void a2x_debug_set(int key, int value);
int a2x_device_info(int* sm_count, int* cc_major, int* cc_minor);
unsigned long long a2x_launch_count(void); /* kernels launched by this library since load */

/* ---------------------------------------------------------------- dense contractions
 * Replace nn.Conv2d / nn.ConvTranspose2d forward + autograd on the path:
 *   opencood/models/common_modules/base_bev_backbone.py:41-105 (blocks: ZeroPad2d(1)+Conv3x3 s2, Conv3x3 p1;
 *   deblocks: ConvTranspose2d k = s), downsample_conv.py:18-32 (shrink 1x1 / 3x3 with bias + ReLU),
 *   airv2x_where2com.py:59-69 (1x1 heads).
 *
 * a2x_conv_shape: n images of h x w pixels (INPUT grid), cin -> cout channels (multiples of 32),
 * ksize in {1,3} with padding ksize/2, stride in {1,2}. For the transposed op (deconv_*), ksize == stride = s
 * in {1,2,4} and the output grid is (h*s) x (w*s).
 */
typedef struct {
    int n, h, w, cin, cout, ksize, stride;
} a2x_conv_shape;

/* Precision. Every GEMM operand v may be given as one plane (`b16 == NULL`: plain TF32; the MMA truncates the fp32
 * bits) or as the 3-term SPLIT:
 *     hi = v (fp32, read by the non-GEMM consumers),   b16 plane 0 = h16 = bf16(v),   b16 plane 1 = l16 = bf16(v - h16)
 * and the contraction is evaluated as  h16*h16 + l16*h16 + h16*l16  (three kind::f16 bf16 MMAs) into one fp32 TMEM
 * accumulator: every product is exact in fp32, the dropped terms are ~2^-17 relative per product (logits within 1e-4 of
 * the fp32 reference through the 23-layer stack) at 3 bf16-MMA passes. Split operands need channel counts that are
 * multiples of 64. */
typedef struct {
    const float* hi;     /* NHWC fp32, pixel stride `cs` elements */
    const void* b16;     /* NHWC bf16 [2 planes], same pixel stride; NULL = single-plane mode */
    long long b16_plane; /* elements between the two bf16 planes */
    int cs;
} a2x_operand;
typedef struct {
    float* hi;
    void* b16;           /* NULL: write the fp32 value only */
    long long b16_plane;
    int cs;
} a2x_output;
typedef struct {
    const float* w32;    /* packed fp32 plane (tf32-rounded; the single-plane operand) */
    const void* w16;     /* packed bf16 planes [2][...] (h16, l16); may be NULL in single-plane mode */
} a2x_weights;

/* weight re-layout (tiny, HBM-bound):  OIHW -> [tap][cout_pad][cin] (forward B operand) and [tap][cin][cout_pad]
 * (dgrad); each as an fp32 plane (tf32_rn) and a bf16 pair [2][...]. Any output pointer may be NULL. */
int a2x_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, float* w_fwd, void* w_fwd16,
                         float* w_dgrad, void* w_dgrad16, a2x_stream_t stream);
/* [tap][cout_pad][cin] -> OIHW (first `cout` rows) ; accumulate != 0 adds into dw_oihw */
int a2x_unpack_conv_wgrad(const float* dw_packed, int cout, int cin, int ksize, int cout_pad, float* dw_oihw,
                          int accumulate, a2x_stream_t stream);
/* ConvTranspose2d weight [cin][cout][s][s] -> [(i*s+j)*cout+co][ci] (forward) and [(i*s+j)][ci][co] (dgrad) */
int a2x_pack_deconv_weight(const float* w_iohw, int cin, int cout, int s, float* w_fwd, void* w_fwd16, float* w_dgrad,
                           void* w_dgrad16, a2x_stream_t stream);
/* [(i*s+j)][ci][co] -> [cin][cout][s][s] */
int a2x_unpack_deconv_wgrad(const float* dw_packed, int cin, int cout, int s, float* dw_iohw, int accumulate,
                            a2x_stream_t stream);

/* Batched forms (one launch per step instead of one per layer): device-resident job tables, <= 128 jobs, elem_begin =
 * running sum of `elems`. kind 0 = conv (a = rows of this source, b = cin, kk = taps; the source's rows land at
 * [row0, row0 + a) of the cout_pad-wide packed tensors, so several parameters can share one fused GEMM operand),
 * kind 1 = deconv (a = cin, b = cout, kk = s*s), kind 2 = plain vector copy (pack: src -> f32 + row0; unpack: double
 * sums[row0 + i] -> float dst[i]). */
typedef struct {
    const float* src;
    float* f32;
    void* f16;
    float* d32;
    void* d16;
    int kind, a, b, kk, cout_pad, row0;
    long long elem_begin, elems;
} a2x_pack_job;
typedef struct {
    const void* src;
    float* dst;
    int kind, a, b, kk, cout_pad, row0;
    long long elem_begin, elems;
} a2x_unpack_job;
int a2x_pack_weights_batched(const a2x_pack_job* jobs_dev, int njobs, long long total_elems, a2x_stream_t stream);
int a2x_unpack_wgrads_batched(const a2x_unpack_job* jobs_dev, int njobs, long long total_elems, a2x_stream_t stream);

/* y = act(scale[c] * conv(x, w) + shift[c]); scale/shift may be NULL.
 * stats != NULL (requires no scale/shift/relu): the epilogue also accumulates the BatchNorm batch statistics of the
 * raw output, stats[c] += sum, stats[cout + c] += sum of squares (doubles, caller zeroes) — no extra HBM pass. */
int a2x_conv2d_fwd(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                   const float* scale, const float* shift, int relu, double* stats, a2x_stream_t stream);
/* Same with act in {0 none, 1 ReLU, 2 GELU(erf)} and accumulate != 0: y = act(scale*conv + shift + y_previous)
 * (residual add in the epilogue). With ksize == 1 this is the token-wise nn.Linear of the transformer fusion
 * networks (cobevt_modules/swap_fusion_modules.py:40-45, base_transformer.py:16-28): rows = n*h*w tokens.
 * ksize == 7 with stride == 2 (padding 3) is the forward-only stem of BevEncode (sub_modules/lss_submodule.py:318):
 * raw fp32 output only (no scale / shift / act / accumulate / stats), weights packed with ksize 7. */
int a2x_conv2d_fwd_ex(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                      const float* scale, const float* shift, int act, int accumulate, double* stats,
                      a2x_stream_t stream);
/* y = residual + dropout(x W^T + bias): the residual sublayers of the transformer fusion networks in train mode
 * (PreNormResidual around Attention.to_out = Linear + Dropout, swap_fusion_modules.py:40-45,127; FeedForward's second
 * Linear + Dropout, cobevt_modules/base_transformer.py:27-41; v2xvit_modules/base_transformer.py:17-46) in the GEMM's
 * epilogue. ksize = stride = 1; y dense fp32 [n][h][w][cout] (y_cs == cout). The mask element of y's element i is
 * elem_offset + i of the site (a2x_dropout_mask exports it): elem_offset != 0 when y is a slice of the site's tensor
 * (the per-type a_linears of HGT, hmsa.py:150-156). residual may be null or alias y; p == 0: no dropout. */
int a2x_linear_dropout_residual_fwd(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, float* y, int y_cs,
                                    const float* bias, const float* residual, unsigned long long seed, unsigned int site,
                                    float p, long long elem_offset, a2x_stream_t stream);
/* dx (+)= conv_transpose(dy, w)   (dx: fp32 NHWC, pixel stride dx_cs) */
int a2x_conv2d_dgrad(const a2x_conv_shape* s, const a2x_operand* dy, const a2x_weights* w, float* dx, int dx_cs,
                     int accumulate, a2x_stream_t stream);
/* Same, with pass 1 of the CONSUMER layer's BatchNorm(train)+ReLU backward fused into the epilogue (3x3 stride 1 only):
 * dx is that layer's dy, so sums[c] += sum g and sums[C + c] += sum g*zhat with g = dx * [z*scale+shift > 0],
 * zhat = (z - mean)*invstd are accumulated while dx is written (a2x_bn_relu_bwd_reduce becomes unnecessary). */
typedef struct {
    const float* z;      /* the consumer layer's pre-BN activation [n][h][w][cin], pixel stride z_cs == dx_cs */
    int z_cs;
    const float* scale;
    const float* shift;
    const float* mean;
    const float* invstd;
    double* sums;        /* [A2X_BN_BWD_REPLICAS][2*cin], zeroed by the caller (copy 0 is written) */
} a2x_bn_bwd_stats;
int a2x_conv2d_dgrad_ex(const a2x_conv_shape* s, const a2x_operand* dy, const a2x_weights* w, float* dx, int dx_cs,
                        int accumulate, const a2x_bn_bwd_stats* bn_stats, a2x_stream_t stream);
/* dw_packed[tap][cout][cin] += sum_pixels dy (x) x   (caller zeroes dw_packed) */
int a2x_conv2d_wgrad(const a2x_conv_shape* s, const a2x_operand* x, const a2x_operand* dy, float* dw_packed,
                     a2x_stream_t stream);

int a2x_deconv_fwd(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                   const float* scale, const float* shift, int relu, double* stats, a2x_stream_t stream);
int a2x_deconv_dgrad(const a2x_conv_shape* s, const a2x_operand* dy, const a2x_weights* w, float* dx, int dx_cs,
                     int accumulate, a2x_stream_t stream);
int a2x_deconv_wgrad(const a2x_conv_shape* s, const a2x_operand* x, const a2x_operand* dy, float* dw_packed,
                     a2x_stream_t stream);

/* ---------------------------------------------------------------- operand split
 * hi = x, b16 = (bf16(x), bf16(x - bf16(x)))  (n multiple of 4; out->cs ignored) */
int a2x_split(const float* x, long long n, const a2x_output* out, a2x_stream_t stream);

/* ---------------------------------------------------------------- transformer fusion (CoBEVT / V2X-ViT) token kernels
 * Token tensors are fp32 NHWC [B*L][H][W][C]. Replace nn.LayerNorm (cobevt_modules/base_transformer.py:6-13),
 * Attention.forward (cobevt_modules/swap_fusion_modules.py:78-127) with its window "(x w1) (y w2)" / grid
 * "(w1 x) (w2 y)" partitions (:155-195), the agent mean + LayerNorm of the mlp head (:268-275) and regroup
 * (cobevt_modules/fuse_utils.py:13-63). */
int a2x_layernorm_fwd(const float* x, int x_cs, long long rows, int C, const float* gamma, const float* beta, float eps,
                      const a2x_output* y, a2x_stream_t stream);
/* y[b][p][:] = LayerNorm(mean_l x[b][l][p][:])  — padded agents are part of the mean, as in the reference */
int a2x_agent_mean_layernorm(const float* x, int B, int L, long long pix, int C, const float* gamma, const float* beta,
                             float eps, const a2x_output* y, a2x_stream_t stream);
/* dst[b][l] = src[scene_start[b] + l] (l < scene_len[b]) else zeros; images of img_elems floats */
int a2x_regroup(const float* src, const int* scene_start, const int* scene_len, int B, int L, long long img_elems,
                const a2x_output* dst, a2x_stream_t stream);
/* Same, the agents' images addressed through a DEVICE table of pointers (agent order), each of which may point into a
 * peer GPU's memory (NVLink P2P): the agent all-gather fused into the regroup (agents one per GPU). */
int a2x_regroup_ptrs(const float* const* src_ptrs_dev, const int* scene_start, const int* scene_len, int B, int L,
                     long long img_elems, const a2x_output* dst, a2x_stream_t stream);
/* qkv: [B*L][H][W][3*heads*dim_head] (q | k | v); bias_table: [(2L-1)(2w-1)^2][heads]; key_mask: int32 [B][L] or NULL;
 * out: dense [B*L][H][W][heads*dim_head]. softmax(q*scale . k + bias) v per (window | grid cell, head). */
int a2x_window_attention_fwd(const float* qkv, const float* bias_table, const int* key_mask, int B, int L, int H, int W,
                             int heads, int dim_head, int window, int grid_mode, float scale, const a2x_output* out,
                             a2x_stream_t stream);

/* Backward of the token kernels (training step of the transformer fusion; autograd of nn.LayerNorm / nn.GELU /
 * Attention.forward). layernorm_bwd: dx_accum += dLN(dy) (the residual stream's gradient), dgamma/dbeta double [C]
 * accumulators (caller zeroes). gelu_fwd / gelu_bwd: erf GELU of a pre-activation kept in fp32. window_attention_bwd:
 * dqkv [B*L][H][W][3*heads*dim_head] from dout (gradient of the attention output), everything recomputed from qkv;
 * dbias_table accumulated with atomics (caller zeroes). */
/* out[b][p] = mean over the L agents of x[b][l][p] (images of img_elems floats) and its adjoint dst[b][l] = scale*src[b] */
int a2x_agent_mean(const float* x, int B, int L, long long img_elems, float* out, a2x_stream_t stream);
int a2x_agent_broadcast(const float* src, int B, int L, long long img_elems, float scale, float* dst, a2x_stream_t stream);
int a2x_layernorm_bwd(const float* x, int x_cs, const float* dy, int dy_cs, long long rows, int C, const float* gamma,
                      float eps, float* dx_accum, int dx_cs, double* dgamma, double* dbeta, a2x_stream_t stream);
int a2x_gelu_fwd(const float* x, long long n, const a2x_output* y, a2x_stream_t stream);
int a2x_gelu_bwd(const float* dy, const float* x, long long n, const a2x_output* dx, a2x_stream_t stream);
/* nn.Dropout of the transformer fusion networks in train mode (opencood/models/cobevt_modules/base_transformer.py:27-56,
 * swap_fusion_modules.py:43, v2xvit_modules/base_transformer.py:17-46, hmsa.py:18,155, mswin.py:47): counter-based
 * (Philox4x32-10) keep flags, a pure function of (seed, site, element index), regenerated by the backward kernels.
 *   a2x_dropout_apply      out = (residual ? residual : 0) + y * keep / (1 - p)      (dense tensors of n elements, n % 8 == 0;
 *                          residual may alias out->hi: the fused "x += dropout(linear(...))" of PreNormResidual; the
 *                          backward calls it on the stream gradient to get d(linear output))
 *   a2x_gelu_dropout_fwd   y  = gelu(x) * keep / (1 - p)                              (FeedForward: Linear, GELU, Dropout)
 *   a2x_gelu_dropout_bwd   dx = dy * keep / (1 - p) * gelu'(x)
 *   a2x_dropout_mask       the keep flags (uint8) of a site: for tests that feed identical masks to the oracle
 * p = 0 turns the mask off (plain copy / GELU). */
int a2x_dropout_apply(const float* y, const float* residual, long long n, unsigned long long seed, unsigned int site, float p,
                      const a2x_output* out, a2x_stream_t stream);
int a2x_gelu_dropout_fwd(const float* x, long long n, unsigned long long seed, unsigned int site, float p, const a2x_output* y,
                         a2x_stream_t stream);
int a2x_gelu_dropout_bwd(const float* dy, const float* x, long long n, unsigned long long seed, unsigned int site, float p,
                         const a2x_output* dx, a2x_stream_t stream);
int a2x_dropout_mask(long long n, unsigned long long seed, unsigned int site, float p, unsigned char* mask, a2x_stream_t stream);
int a2x_window_attention_bwd(const float* qkv, const float* dout, const float* bias_table, const int* key_mask, int B,
                             int L, int H, int W, int heads, int dim_head, int window, int grid_mode, float scale,
                             float* dqkv, float* dbias_table, a2x_stream_t stream);
/* Same, with the gradient written as a GEMM operand: dqkv->hi (fp32, may be NULL) and / or the bf16 split planes — its
 * consumers are the to_qkv weight / data gradient GEMMs, so no separate conversion pass over the [.., 3*D] tensor. */
int a2x_window_attention_bwd_split(const float* qkv, const float* dout, const float* bias_table, const int* key_mask, int B,
                                   int L, int H, int W, int heads, int dim_head, int window, int grid_mode, float scale,
                                   const a2x_output* dqkv, float* dbias_table, a2x_stream_t stream);

/* ---------------------------------------------------------------- V2X-ViT fusion (csrc/v2xvit.cu)
 * x[a][p][:] += Linear(emb_table[emb_idx[a]])  — RTE, v2xvit_modules/v2xvit_basic.py:41-80. vec_ws: [n_agents][C]. */
int a2x_rte_add(float* x, int n_agents, long long pix, int C, const float* emb_table, const int* emb_idx_dev,
                const float* lin_w, const float* lin_b, float* vec_ws, a2x_stream_t stream);
/* Fold relation_att / relation_msg ([4][heads][dh][dh]) into the typed K / V projections (hmsa.py:37-151): for key type
 * tj the fused projection w_fold[tj] is [5C][C] = q | k' (query type 0) | k' (query type 1) | v' (0) | v' (1), with
 * b_fold[tj] [5C]. qw..vb: HOST arrays of two device pointers (agent types 0 / 1). */
int a2x_hgt_fold(const float* const* qw, const float* const* qb, const float* const* kw, const float* const* kb,
                 const float* const* vw, const float* const* vb, const float* relation_att, const float* relation_msg,
                 int C, int heads, float* w_fold, float* b_fold, a2x_stream_t stream);
/* per pixel, per head: softmax over the agents j valid at that pixel (key_mask[j][p] != 0) of scale * q_i . k'_{type_i}(j),
 * out_i = sum_j att v'_{type_i}(j). qkv: [n][pix][5C] from the folded projection; out: dense [n][pix][C]. */
int a2x_hgt_attention_fwd(const float* qkv, const int* types_dev, const float* key_mask, int n_agents, long long pix,
                          int heads, int dim_head, float scale, const a2x_output* out, a2x_stream_t stream);
/* SplitAttn over the three pyramid-window branches + residual (split_attn.py:28-63):
 * x += sum_r softmax_r(fc2 relu(LN(fc1 mean_p(w0+w1+w2))))[r] * w_r. sums_ws: [n][C] (written: the pooled sums, kept for
 * the backward), partials_ws: [n][A2X_SPLIT_ATTN_CHUNKS][C] scratch (per-chunk sums reduced in a fixed order: the result
 * is bit-reproducible and independent of how many agents share the call), weights_ws: [n][3][C]. */
#define A2X_SPLIT_ATTN_CHUNKS 64
int a2x_split_attn_fuse(const float* w0, const float* w1, const float* w2, int n_agents, long long pix, int C,
                        const float* fc1, const float* ln_gamma, const float* ln_beta, const float* fc2,
                        float* sums_ws, float* partials_ws, float* weights_ws, float* x_inout, a2x_stream_t stream);

/* Backward of the V2X-ViT kernels. hgt_attention_bwd: dqkv [n][pix][5C] (every slot written) from dout [n][pix][C].
 * hgt_fold_bwd: gradients of the fused projection (dw_fold [2][5C][C], db_fold [2][5C]) -> typed q/k/v linears (written)
 * and relation_att / relation_msg (written). split_attn_bwd: d(win_r) as three split operands, gradients of fc1 / bn1 /
 * fc2 accumulated with atomics (caller zeroes); sums_saved / weights_saved are the forward's [n][C] pooled sums and
 * [n][3][C] mixing weights. rte_bwd: dvec_sums = per-agent column sums of dx ([n][2C] doubles from a2x_channel_stats);
 * dlin_w / dlin_b / demb_table accumulated with atomics (caller zeroes). */
int a2x_hgt_attention_bwd(const float* qkv, const int* types_dev, const float* key_mask, const float* dout, int n_agents,
                          long long pix, int heads, int dim_head, float scale, float* dqkv, a2x_stream_t stream);
int a2x_hgt_fold_bwd(const float* dw_fold, const float* db_fold, const float* const* kw, const float* const* kb,
                     const float* const* vw, const float* const* vb, const float* relation_att, const float* relation_msg,
                     int C, int heads, float* const* dqw, float* const* dqb, float* const* dkw, float* const* dkb,
                     float* const* dvw, float* const* dvb, float* drelation_att, float* drelation_msg, a2x_stream_t stream);
int a2x_split_attn_bwd(const float* dx, const float* w0, const float* w1, const float* w2, int n_agents, long long pix, int C,
                       const float* fc1, const float* ln_gamma, const float* ln_beta, const float* fc2, const float* sums_saved,
                       const float* weights_saved, float* dw_ws, float* dgap_ws, const a2x_output* d0, const a2x_output* d1,
                       const a2x_output* d2, float* dfc1, float* dln_gamma, float* dln_beta, float* dfc2,
                       a2x_stream_t stream);
int a2x_rte_bwd(const double* dvec_sums, int n_agents, int C, const float* emb_table, const int* emb_idx_dev, const float* lin_w,
                float* dlin_w, float* dlin_b, float* demb_table, a2x_stream_t stream);

/* ---------------------------------------------------------------- ego-warp (affine bilinear resampling, NHWC)
 * warp_affine_simple = F.affine_grid + F.grid_sample(bilinear | nearest, zeros padding)
 * (common_modules/torch_transformation_utils.py:327-334); theta: [n][2][3] normalised matrices. */
int a2x_warp_affine_fwd(const float* src, int src_cs, const float* theta, int n, int hi, int wi, int c, int ho, int wo,
                        int align_corners, int nearest, const a2x_output* dst, a2x_stream_t stream);
/* out[a][p] = valid[a] (NULL = all valid) && nearest source pixel of p inside the map: the rotated ROI mask of
 * get_roi_and_cav_mask (torch_transformation_utils.py:15-113) */
int a2x_roi_mask(const float* theta, const int* valid, int n, int h, int w, int align_corners, float* out,
                 a2x_stream_t stream);
/* dsrc += bilinear^T dout (caller zeroes dsrc) */
int a2x_warp_affine_bwd(const float* dout, int dout_cs, const float* theta, int n, int hi, int wi, int c, int ho, int wo,
                        int align_corners, float* dsrc, int dsrc_cs, a2x_stream_t stream);

/* ---------------------------------------------------------------- detection decode + rotated NMS (batch 1 inference)
 * Replaces VoxelPostprocessor.post_process_airv2x (data_utils/post_processor/voxel_postprocessor.py:666-840):
 * objectness gate (sigmoid(obj) > obj_threshold), delta_to_boxes3d (:585-635), boxes_to_corners_3d order "hwl"
 * (utils/box_utils.py:195-258), remove_large_pred_bbx / remove_bbx_abnormal_z (:981-1035), nms_rotated (:823-868:
 * top-1000 by score, greedy, IoU of the first four corners' polygons > nms_threshold; shapely -> convex-quad clipping
 * in double precision), range mask (:399-430). heads: NHWC [H][W][heads_cs] = cls (class-major, A*num_class) | reg
 * (7A) | obj (A); anchors: [H][W][A][7] (x, y, z, h, w, l, yaw). Outputs (device, score-descending): corners
 * [max_out][8][3], scores, labels, boxes [max_out][7], anchor index; *n_out_dev = count. *status_dev != 0 reports an
 * internal capacity overflow (bit 0: candidates, bit 1: score ties at the top-1000 cut). No host sync. */
size_t a2x_postprocess_workspace_bytes(int n_anchors);
int a2x_postprocess_det(const float* heads, int heads_cs, int H, int W, int A, int num_class, const float* anchors,
                        float obj_threshold, float nms_threshold, const float* lidar_range6, void* workspace,
                        size_t workspace_bytes, float* out_corners, float* out_scores, int* out_labels, float* out_boxes,
                        int* out_anchor_index, int max_out, int* n_out_dev, int* status_dev, a2x_stream_t stream);
/* out[i][j] = IoU of the xy polygons (first four corners) of boxes_a[i] and boxes_b[j], each [n][8][3]: the matching
 * step of caluclate_tp_fp (utils/eval_utils_opv2v.py:41-95) */
int a2x_rotated_iou_matrix(const float* boxes_a, int na, const float* boxes_b, int nb, float* out, a2x_stream_t stream);

/* ---------------------------------------------------------------- BatchNorm / ReLU / masks (HBM-bound)
 * Replace nn.BatchNorm2d(eps 1e-3, momentum 0.01) + nn.ReLU and their autograd
 * (opencood/models/common_modules/base_bev_backbone.py:52-66, :82-90) and the bias+ReLU of
 * downsample_conv.py:18-32. x/y/z/dy are NHWC with pixel strides *_cs; npix = n*h*w; C multiple of 4.
 */
/* sums[c] += sum x, sums[C+c] += sum x^2 (double; caller zeroes) */
int a2x_channel_stats(const float* x, int x_cs, long long npix, int C, double* sums, a2x_stream_t stream);
/* batch stats -> scale = gamma*invstd, shift = beta - mean*scale; running stats updated n_updates times */
int a2x_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps, float momentum,
                    int n_updates, float* running_mean, float* running_var, int C, float* scale, float* shift,
                    float* mean_out, float* invstd_out, a2x_stream_t stream);
int a2x_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int C, float* scale, float* shift, a2x_stream_t stream);
/* y = relu?(x*scale[c] + shift[c]) * mask[pixel]   (scale/shift/mask may be NULL; y->b16 != NULL -> split) */
int a2x_affine_act(const float* x, int x_cs, const float* scale, const float* shift, int relu, const float* mask,
                   const a2x_output* y, long long npix, int C, a2x_stream_t stream);
/* bn_finalize + affine_act in one pass (train mode): y = relu?(BN_batch(z)); publishes scale / shift / mean / invstd
 * (read by the backward kernels) and applies `n_updates` running-statistics updates. sums = [sum, sum of squares]. */
int a2x_bn_train_act(const float* z, int z_cs, const double* sums, double count, const float* gamma, const float* beta,
                     float eps, float momentum, int n_updates, float* running_mean, float* running_var, float* scale,
                     float* shift, float* mean_out, float* invstd_out, int relu, const a2x_output* y, long long npix,
                     int C, a2x_stream_t stream);
/* BN(train)+ReLU backward, pass 1: sum g and sum g*zhat with g = dy*(z*scale+shift > 0), accumulated into
 * A2X_BN_BWD_REPLICAS copies of [2C] doubles (`sums` = [replicas][2C] + one ticket word, all zeroed by the caller; block b
 * adds into copy b % replicas, which spreads the same-address atomics, and the last block folds the copies into copy 0,
 * which is what pass 2 reads). Allocate A2X_BN_BWD_SUMS(C) doubles. */
#define A2X_BN_BWD_SUMS(C) (2 * (C) * A2X_BN_BWD_REPLICAS + 1)
#define A2X_BN_BWD_REPLICAS 8
int a2x_bn_relu_bwd_reduce(const float* dy, int dy_cs, const float* z, int z_cs, const float* scale, const float* shift,
                           const float* mean, const float* invstd, long long npix, int C, double* sums,
                           a2x_stream_t stream);
/* pass 2: dz = scale*(g - sum_g/count - zhat*sum_gz/count); dgamma/dbeta (may be NULL) from the sums */
int a2x_bn_relu_bwd_apply(const float* dy, int dy_cs, const float* z, int z_cs, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const double* sums, double count,
                          const a2x_output* dz, long long npix, int C, float* dgamma, float* dbeta,
                          int accumulate_param_grads, a2x_stream_t stream);
/* g = dy * (y > 0) * mask[pixel]  (y, mask may be NULL) */
int a2x_relu_bwd(const float* dy, int dy_cs, const float* y, int y_cs, const float* mask, const a2x_output* g,
                 long long npix, int C, a2x_stream_t stream);
/* out[c] (+)= (float) sums[c] : bias gradients from a2x_channel_stats sums */
int a2x_sums_to_float(const double* sums, int C, float* out, int accumulate, a2x_stream_t stream);
/* torch.count_nonzero (airv2x_where2com.py:122) */
int a2x_count_nonzero(const float* x, long long n, unsigned long long* out, a2x_stream_t stream);

/* ---------------------------------------------------------------- Lift-Splat camera branch (lift + voxel pooling)
 * Replace the lift `depth.unsqueeze(1) * x_img.unsqueeze(2)` of CamEncode.forward
 * (opencood/models/sub_modules/lss_submodule.py:170-186) together with LiftSplatShootEncoder.voxel_pooling
 * (opencood/models/common_modules/airv2x_encoder.py:208-275) without materialising the [B,N,D,fH,fW,C] product.
 * depth [B*N][D][fH][fW] (softmax over D), feat [B*N][C][fH][fW], geom [B][N][D][fH][fW][3] (get_geometry, :133-168);
 * origin3 = bx - dx / 2, dx3, nx3 = gen_dx_bx (utils/camera_utils.py:238-244). bev: NHWC [B][ny][nx][nz*C] (channel
 * z*C + c = the reference's [B, C*nz, ny, nx] after its unbind/cat), zero-filled here. cells_ws [B*N*D*fH*fW] i32 keeps
 * every frustum point's cell (-1 = outside) for the backward: ddepth / dfeat in the layouts of depth / feat. */
int a2x_lift_splat_fwd(const float* depth, const float* feat, const float* geom, int B, int N, int D, int fH, int fW, int C,
                       const float* origin3, const float* dx3, const int* nx3, float* bev, int* cells_ws, a2x_stream_t stream);
int a2x_lift_splat_bwd(const float* depth, const float* feat, const int* cells_ws, const float* dbev, int B, int N, int D,
                       int fH, int fW, int C, float* ddepth, float* dfeat, a2x_stream_t stream);

/* ---------------------------------------------------------------- anchor-target assignment (label generation)
 * Replaces VoxelPostprocessor.generate_label_airv2x + collate_batch_airv2x
 * (opencood/data_utils/post_processor/voxel_postprocessor.py:217-354, :392-430) and bbox_overlaps
 * (opencood/utils/box_overlaps.pyx:17-56) for a batch of B samples. anchor_standup [N][4] f32 / gt_standup [total_gt][4]
 * f32: axis-aligned (xmin, ymin, xmax, ymax) footprints of the rotated boxes (host: labels.py, the reference's fp32
 * corner arithmetic); anchors [N][7] f64 (x,y,z,h,w,l,yaw), N = H*W*A in (h, w, a) order; gt_boxes [total_gt][7] f64 and
 * gt_class [total_gt]: the valid boxes of every sample back to back, gt_offsets_dev [B+1]. Outputs in the layout
 * a2x_det_loss reads: targets [B][N][7] f32, pos_equal_one / neg_equal_one [B][N] f32, class_ids [B][N] i32.
 * Workspaces: code_ws [B][N] i32, best_ws [max(total_gt,1)] u64. */
int a2x_assign_targets(const float* anchor_standup, const double* anchors, int n_anchors, const float* gt_standup,
                       const double* gt_boxes, const int* gt_class, const int* gt_offsets_dev, int total_gt, int B,
                       float pos_threshold, float neg_threshold, int* code_ws, unsigned long long* best_ws, float* targets,
                       float* pos_equal_one, float* neg_equal_one, int* class_ids, a2x_stream_t stream);

/* ---------------------------------------------------------------- voxelisation
 * Replaces SpVoxelPreprocessor.preprocess -> spconv Point2VoxelCPU3d.point_to_voxel + collate_batch
 * (opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:59-72, :96-116, :142-175), bit-exact with the
 * sequential first-come algorithm. points: [total][4] f32 of all agents concatenated, offsets_dev[n_agents+1]
 * (device). Outputs are fixed-capacity slabs per agent (cap >= max_voxels): voxels [n][cap][32][4],
 * coords [n][cap][4] (agent,z,y,x), num_points [n][cap], counts [n] (voxels per agent, device).
 */
size_t a2x_voxelize_workspace_bytes(int n_agents, long long total_points, int nx, int ny, int nz, int cap);
/* ego_flags (device u8 per agent, may be NULL) / strict_range fold the dataset's point filters in
 * (mask_ego_points / mask_points_by_range, opencood/utils/pcd_utils.py:136-190): dropped points keep their place in
 * the input order, so the first-come semantics are unchanged. */
int a2x_voxelize(const float* points, const int* offsets_dev, int n_agents, long long total_points, const float* range6,
                 const float* vsize3, int max_points, int max_voxels, int cap, const unsigned char* ego_flags,
                 int strict_range, void* workspace, size_t workspace_bytes, float* voxels, int* coords, int* num_points,
                 int* counts, a2x_stream_t stream);
/* Same, with the dataset's agent -> ego projection folded in (project_points_by_matrix_torch, utils/box_utils.py:1038-1066,
 * called between mask_ego_points and mask_points_by_range at
 * data_utils/datasets/airv2x/intermediate_fusion_dataset.py:592-600): transforms_dev [n_agents][4][4] f32 row-major
 * (device, may be NULL = identity, no arithmetic). Points arrive in each agent's sensor frame; the ego-box test uses the
 * sensor-frame coordinates, the range test / voxel index / stored pillar points the projected ones, bit-exact with
 * torch's fp32 evaluation of the reference's einsum. */
int a2x_voxelize_ex(const float* points, const int* offsets_dev, const float* transforms_dev, int n_agents,
                    long long total_points, const float* range6, const float* vsize3, int max_points, int max_voxels,
                    int cap, const unsigned char* ego_flags, int strict_range, void* workspace, size_t workspace_bytes,
                    float* voxels, int* coords, int* num_points, int* counts, a2x_stream_t stream);

/* ---------------------------------------------------------------- PillarVFE + PointPillarScatter
 * Replace PillarVFE.forward / PFNLayer.forward / PointPillarScatter.forward
 * (opencood/models/common_modules/airv2x_pillar_vfe.py:105-160, :27-49; point_pillar_scatter.py:15-82).
 * voxels [M][32][4] f32 zero padded, num_points [M] i32, coords [M][4] i32 (agent,z,y,x) — the dict
 * SpVoxelPreprocessor.collate_batch emits (data_utils/pre_processor/sp_voxel_preprocessor.py:142-175).
 * agent_map[agent] = row of that agent in the scene-major canvas [n_total][ny][nx][64].
 */
typedef struct {
    float voxel_x, voxel_y, voxel_z;    /* per agent type (airv2x_pillar_vfe.py:84-89) */
    float x_offset, y_offset, z_offset; /* voxel/2 + range_lo */
    int nx, ny;
} a2x_pfn_geom;
/* Optional segmented addressing (voxeliser slabs, no host sync): flat pillar p in [0, nseg*seg_cap) lives in slab
 * seg_ids[p / seg_cap] (NULL ids = identity) at index p % seg_cap and is valid while < seg_counts[slab] (device
 * array). Pass seg == NULL for the reference dict layout (plain [0, m)). With segments, `m` / `rows` are ignored. */
typedef struct {
    const int* seg_ids;
    const int* seg_counts;
    int seg_cap, nseg;
} a2x_pfn_segments;
/* moments65 = [sum f (10), upper triangle of sum f f^T (55)] over all M*32 rows (zeroed here) */
int a2x_pfn_moments(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                    const a2x_pfn_segments* seg, double* moments65, a2x_stream_t stream);
/* BatchNorm1d batch statistics of W f from the moments (rows = M*32) */
int a2x_pfn_stats_finalize(const double* moments65, double rows, const a2x_pfn_segments* seg, const float* w,
                           const float* gamma, const float* beta,
                           float eps, float momentum, int n_updates, float* running_mean, float* running_var,
                           float* scale, float* shift, float* mean_out, float* invstd_out, a2x_stream_t stream);
/* canvas[agent_map[a]][y][x][:] = max_slot relu(scale*(W f)+shift); optional pillar_out [M][64], amax [M][64] u8 */
int a2x_pfn_scatter(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                    const a2x_pfn_segments* seg, const float* w, const float* scale, const float* shift,
                    const int* agent_map, const a2x_output* canvas, float* pillar_out, unsigned char* amax,
                    a2x_stream_t stream);
/* same, with (i) canvas->hi optional (null: only the bf16 split planes the block-0 tap-GEMM reads are written) and
 * (ii) *nonzero_count += number of non-zero canvas values written: `spatial_features.count_nonzero()` of
 * opencood/models/airv2x_where2com.py:122 without a pass over the canvas (caller zeroes the counter) */
int a2x_pfn_scatter_ex(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                       const a2x_pfn_segments* seg, const float* w, const float* scale, const float* shift,
                       const int* agent_map, const a2x_output* canvas, float* pillar_out, unsigned char* amax,
                       long long* nonzero_count, a2x_stream_t stream);
/* train-mode backward to (W, gamma, beta) given d(canvas); acc_ws = 64*12 doubles of workspace */
int a2x_pfn_bwd(const float* voxels, const int* num_points, const int* coords, long long m, const a2x_pfn_geom* geom,
                const a2x_pfn_segments* seg, const float* w, const float* scale, const float* shift, const float* mean, const float* invstd,
                const int* agent_map, const float* dcanvas, const unsigned char* amax, const double* moments65,
                double rows, double* acc_ws, float* dw, float* dgamma, float* dbeta, int accumulate,
                a2x_stream_t stream);

/* ---------------------------------------------------------------- Where2comm communication + fusion
 * Replace Communication.forward (opencood/models/where2comm_modules/where2comm_fuse.py:83-149) and
 * AttentionFusion.forward (:152-164, :14-45).
 */
/* conf[p] = max_{c<ncls} sigmoid(psm[p][c]) */
int a2x_comm_confidence(const float* psm, int psm_cs, int ncls, long long npix, float* conf, a2x_stream_t stream);
/* smooth = gaussian_filter(conf) (ksz x ksz weights + bias; ksz = 0 -> identity); write_mask: mask = smooth > thr */
int a2x_comm_smooth_mask(const float* conf, const float* gauss_w, const float* gauss_b, int ksz, int n, int h, int w,
                         float threshold, int write_mask, float* smooth, float* mask, a2x_stream_t stream);
/* train mode: mask[a] = 1 on the k_per_agent[a] largest smooth values of agent a (device array) */
int a2x_comm_topk_mask(const float* smooth, int n, int hw, const int* k_per_agent, float* mask, a2x_stream_t stream);
/* dst [n][H][W] = F.interpolate(src [n][h][w], bilinear, align_corners=False): the mask resize of
 * where2comm_fuse.py:230-236 (legacy stride-2 shrink header) */
int a2x_resize_bilinear(const float* src, int n, int h, int w, float* dst, int H, int W, a2x_stream_t stream);
/* Sparse feature select (the payload an agent transmits, where2comm_fuse.py:237): warp-ballot compaction of the cells
 * of one agent's level-0 map [hw][C] selected by its mask (all cells if force_all, the ego). hdr[0] = records written,
 * hdr[1] = cells the mask itself selected (communication-rate numerator); idx[r] = cell, vals[r][:] = feature row. */
int a2x_mask_compact(const float* x, int x_cs, const float* mask, int force_all, int hw, int C, int* hdr, int* idx,
                     float* vals, a2x_stream_t stream);
/* Receiver: dst [n_agents][hw][C] is zero-filled, then for every agent a the records at bufs_dev[a] (+0: count,
 * +off_idx_bytes: idx, +off_vals_bytes: vals) are scattered into dst[a]. bufs_dev: DEVICE table of per-agent base
 * pointers, local (after an all-gather) or peer-GPU memory (NVLink P2P: the gather fused into the decompaction). */
int a2x_mask_decompact_ptrs(const void* const* bufs_dev, long long off_idx_bytes, long long off_vals_bytes, int n_agents,
                            int hw, int C, float* dst, a2x_stream_t stream);
/* ones[b] = sum of the scene's mask (before ego override); then mask[ego of scene b] = 1 */
int a2x_comm_rate_ego(float* mask, int hw, int n_scenes, const int* scene_start, const int* scene_len, float* ones,
                      a2x_stream_t stream);
/* x: [n_agents][hw][c] of ONE scene (agent 0 = ego) -> out [hw][c] = row 0 of softmax(x x^T / sqrt(c)) x */
int a2x_att_fuse_fwd(const float* x, int n_agents, int hw, int c, const a2x_output* out, a2x_stream_t stream);
int a2x_att_fuse_bwd(const float* x, const float* dout, int n_agents, int hw, int c, float* dx, a2x_stream_t stream);

/* ---------------------------------------------------------------- detection loss
 * Replace PointPillarLossMultiClass.forward (opencood/loss/point_pillar_loss_multiclass.py:96-215, :273-289):
 * value (reg, cls, obj terms) and gradient w.r.t. the NHWC head logits [psm(A*K) | rm(7A) | obj(A)] in one pass.
 */
int a2x_det_loss(const float* heads, int heads_cs, int B, long long HW, int A, int K, const float* targets,
                 const float* pos_equal_one, const int* class_ids, float cls_weight, float reg_coe, float* npos_ws,
                 float* dheads, int dheads_cs, double* loss3, a2x_stream_t stream);
/* Replace PointPillarLoss.forward of the legacy `point_pillar_*` models (opencood/loss/point_pillar_loss.py:77-215): one
 * logit per anchor, focal term / B once, the same smooth-L1 with sin-difference, no objectness; heads [psm(A) | rm(7A)],
 * loss3 = (reg, conf, 0). */
int a2x_det_loss_legacy(const float* heads, int heads_cs, int B, long long HW, int A, const float* targets,
                        const float* pos_equal_one, float cls_weight, float reg_coe, float* npos_ws, float* dheads,
                        int dheads_cs, double* loss3, a2x_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AIRV2X_B200_H */
