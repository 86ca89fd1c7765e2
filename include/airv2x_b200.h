/*
 * airv2x_b200.h — C ABI of the B200-native collaborative-perception hot path.
 *
 * Every entry point is a stream-ordered launcher over raw DEVICE pointers (plain pointers and sizes, no torch
 * types), returns 0 on success (1 = bad argument, 2 = CUDA error; message via a2x_last_error(), thread-local),
 * performs no hidden allocation (callers pass outputs and workspaces) and never synchronises the device.
 * Activations are fp32 NHWC ("pixel-major": [n][h][w][c], c contiguous) with an explicit pixel stride
 * (`*_cs`, elements) so that channel slices of a wider buffer (the 384-channel concat) can be read/written in
 * place. Dense contractions run as tcgen05.mma.kind::tf32 (fp32 storage, fp32 accumulate).
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
void a2x_debug_set(int key, int value);
int a2x_device_info(int* sm_count, int* cc_major, int* cc_minor);

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

/* weight re-layout (tiny, HBM-bound):  OIHW -> [tap][cout][cin] (forward B operand) and [tap][cin][cout] (dgrad) */
int a2x_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, float* w_fwd, float* w_dgrad,
                         a2x_stream_t stream);
/* [tap][cout_pad][cin] -> OIHW (first `cout` rows) ; accumulate != 0 adds into dw_oihw */
int a2x_unpack_conv_wgrad(const float* dw_packed, int cout, int cin, int ksize, int cout_pad, float* dw_oihw,
                          int accumulate, a2x_stream_t stream);
/* ConvTranspose2d weight [cin][cout][s][s] -> [(i*s+j)*cout+co][ci] (forward) and [(i*s+j)][ci][co] (dgrad) */
int a2x_pack_deconv_weight(const float* w_iohw, int cin, int cout, int s, float* w_fwd, float* w_dgrad,
                           a2x_stream_t stream);
/* [(i*s+j)][ci][co] -> [cin][cout][s][s] */
int a2x_unpack_deconv_wgrad(const float* dw_packed, int cin, int cout, int s, float* dw_iohw, int accumulate,
                            a2x_stream_t stream);

/* y = act(scale[c] * conv(x, w) + shift[c]); scale/shift may be NULL; y has pixel stride y_cs */
int a2x_conv2d_fwd(const a2x_conv_shape* s, const float* x, int x_cs, const float* w_fwd, float* y, int y_cs,
                   const float* scale, const float* shift, int relu, a2x_stream_t stream);
/* dx (+)= conv_transpose(dy, w) */
int a2x_conv2d_dgrad(const a2x_conv_shape* s, const float* dy, int dy_cs, const float* w_dgrad, float* dx, int dx_cs,
                     int accumulate, a2x_stream_t stream);
/* dw_packed[tap][cout][cin] += sum_pixels dy (x) x   (caller zeroes dw_packed) */
int a2x_conv2d_wgrad(const a2x_conv_shape* s, const float* x, int x_cs, const float* dy, int dy_cs, float* dw_packed,
                     a2x_stream_t stream);

int a2x_deconv_fwd(const a2x_conv_shape* s, const float* x, int x_cs, const float* w_fwd, float* y, int y_cs,
                   const float* scale, const float* shift, int relu, a2x_stream_t stream);
int a2x_deconv_dgrad(const a2x_conv_shape* s, const float* dy, int dy_cs, const float* w_dgrad, float* dx, int dx_cs,
                     int accumulate, a2x_stream_t stream);
int a2x_deconv_wgrad(const a2x_conv_shape* s, const float* x, int x_cs, const float* dy, int dy_cs, float* dw_packed,
                     a2x_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AIRV2X_B200_H */
