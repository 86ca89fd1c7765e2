// C-ABI launchers for the dense contractions (conv / transposed conv: forward, data gradient, weight gradient).
// All of them lower onto the two tcgen05 kernels in tapgemm.cuh / wgrad.cuh by choosing tensor-map views and taps.
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "tapgemm.cuh"
#include "wgrad.cuh"

namespace a2x {

extern int g_debug[16];

// ------------------------------------------------------------------ tensor-map views of an NHWC tensor
// Rank-5 view (c, w, x, h, n) of every `step`-th pixel starting at (h_off, w_off). box = (32, box_w, 1, box_h, 1).
static int make_act_map(CUtensorMap* m, const float* base, int n, int h, int w, int c, int cs, int step, int h_off,
                        int w_off, int box_w, int box_h, int atom32 = 0) {
    const int hp = (h - h_off + step - 1) / step;
    const int wp = (w - w_off + step - 1) / step;
    if (hp <= 0 || wp <= 0) {
        set_error("empty activation view");
        return 1;
    }
    const float* b = base + ((long long)h_off * w + w_off) * cs;
    uint64_t dims[5] = {(uint64_t)c, (uint64_t)wp, 1, (uint64_t)hp, (uint64_t)n};
    uint64_t str[4] = {(uint64_t)step * cs * 4, (uint64_t)step * cs * 4, (uint64_t)step * w * cs * 4,
                       (uint64_t)h * w * cs * 4};
    uint32_t box[5] = {32, (uint32_t)box_w, 1, (uint32_t)box_h, 1};
    return encode_tmap_f32(m, b, 5, dims, str, box, atom32);
}

// weights [taps][rows][k] (k contiguous); box = (32, bn, 1)
static int make_w_map(CUtensorMap* m, const float* base, int taps, int rows, int k, int bn) {
    uint64_t dims[3] = {(uint64_t)k, (uint64_t)rows, (uint64_t)taps};
    uint64_t str[2] = {(uint64_t)k * 4, (uint64_t)rows * k * 4};
    uint32_t box[3] = {32, (uint32_t)(bn < rows ? bn : rows), 1};
    return encode_tmap_f32(m, base, 3, dims, str, box);
}

static int pick_tw_log2(int gh, int gw, int pix, int lo, int hi) {
    long long best = -1;
    int best_l = lo;
    for (int l = hi; l >= lo; --l) {
        const int tw = 1 << l, th = pix >> l;
        if (th < 1) continue;
        const long long cost = (long long)((gh + th - 1) / th) * th * ((gw + tw - 1) / tw) * tw;
        if (best < 0 || cost < best) {
            best = cost;
            best_l = l;
        }
    }
    return best_l;
}

template <int BN, int STAGES>
static int launch_tg(const TgParams& p, int n_col_tiles, cudaStream_t st) {
    using L = TgSmem<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            L::TOTAL));
        configured = true;
    }
    dim3 grid(p.n_img * p.tiles_h * p.tiles_w, n_col_tiles, 1);
    tapgemm_kernel<BN, STAGES><<<grid, 192, L::TOTAL, st>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Fill grid/tile fields and dispatch on the column-tile width.
static int run_tg(TgParams& p, int gh, int gw, int n_img, int ncols, cudaStream_t st) {
    p.n_img = n_img;
    p.gh = gh;
    p.gw = gw;
    p.ncols = ncols;
    const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
    p.tiles_h = (gh + TH - 1) / TH;
    p.tiles_w = (gw + TW - 1) / TW;
    int bn = ncols >= 256 ? 256 : ncols >= 128 ? 128 : ncols >= 64 ? 64 : 32;
    if (g_debug[1] > 0) bn = g_debug[1];
    if (ncols % 32 != 0) {
        set_error("column count %d not a multiple of 32", ncols);
        return 1;
    }
    const int tiles_n = (ncols + bn - 1) / bn;
    switch (bn) {
        case 256: return launch_tg<256, 4>(p, tiles_n, st);
        case 128: return launch_tg<128, 3>(p, tiles_n, st);
        case 64: return launch_tg<64, 4>(p, tiles_n, st);
        case 32: return launch_tg<32, 4>(p, tiles_n, st);
    }
    set_error("bad bn %d", bn);
    return 1;
}

static int bn_for(int ncols) {
    int bn = ncols >= 256 ? 256 : ncols >= 128 ? 128 : ncols >= 64 ? 64 : 32;
    if (g_debug[1] > 0) bn = g_debug[1];
    return bn;
}

static void set_plain_out(TgParams& p, float* out, int h, int w, int cs, int step, int h_off, int w_off) {
    p.out = out + ((long long)h_off * w + w_off) * cs;
    p.osn = (long long)h * w * cs;
    p.osh = (long long)step * w * cs;
    p.osw = (long long)step * cs;
    p.sub_c = 1 << 30;
    p.sub_s = 1;
    p.sub_sh = 0;
    p.sub_sw = 0;
}

static int check_shape(const a2x_conv_shape* s, bool transposed) {
    if (!s) {
        set_error("null shape");
        return 1;
    }
    if (s->n <= 0 || s->h <= 0 || s->w <= 0 || s->cin <= 0 || s->cout <= 0 || s->cin % 32 || s->cout % 32) {
        set_error("bad conv shape n=%d h=%d w=%d cin=%d cout=%d (channels must be multiples of 32)", s->n, s->h, s->w,
                  s->cin, s->cout);
        return 1;
    }
    if (transposed) {
        if (s->ksize != s->stride || !(s->stride == 1 || s->stride == 2 || s->stride == 4)) {
            set_error("deconv needs ksize == stride in {1,2,4}, got k=%d s=%d", s->ksize, s->stride);
            return 1;
        }
    } else {
        if (!((s->ksize == 1 && s->stride == 1) || (s->ksize == 3 && (s->stride == 1 || s->stride == 2)))) {
            set_error("conv needs (k=1,s=1) or (k=3,s in {1,2}), got k=%d s=%d", s->ksize, s->stride);
            return 1;
        }
    }
    return 0;
}

// forward taps of a k x k conv with padding k/2 and stride s over the input parity maps
static int build_fwd_taps(const a2x_conv_shape* s, TgTap* taps) {
    int nt = 0;
    if (s->ksize == 1) {
        taps[nt++] = TgTap{0, 0, 0, 0, 0, 0};
        return nt;
    }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            TgTap t{};
            if (s->stride == 1) {
                t.map = 0;
                t.dh = (int16_t)(r - 1);
                t.dw = (int16_t)(c - 1);
            } else {
                const int hp = (r == 1) ? 0 : 1, wp = (c == 1) ? 0 : 1;
                t.map = (int16_t)(hp * 2 + wp);
                t.dh = (int16_t)(r == 0 ? -1 : 0);
                t.dw = (int16_t)(c == 0 ? -1 : 0);
            }
            t.dx = 0;
            t.btap = r * 3 + c;
            taps[nt++] = t;
        }
    return nt;
}

static int build_fwd_maps(const a2x_conv_shape* s, const float* x, int x_cs, CUtensorMap* maps, int box_w, int box_h,
                          int atom32 = 0) {
    if (s->stride == 1)
        return make_act_map(&maps[0], x, s->n, s->h, s->w, s->cin, x_cs, 1, 0, 0, box_w, box_h, atom32);
    for (int hp = 0; hp < 2; ++hp)
        for (int wp = 0; wp < 2; ++wp) {
            if (hp >= s->h || wp >= s->w) continue;
            int r = make_act_map(&maps[hp * 2 + wp], x, s->n, s->h, s->w, s->cin, x_cs, 2, hp, wp, box_w, box_h, atom32);
            if (r) return r;
        }
    return 0;
}

template <int BN, int TPC, int STAGES, bool SPLIT>
static int launch_wg(const WgParams& p, int ksplit, int tiles_a, int tap_groups, cudaStream_t st) {
    using L = WgSmem<BN, TPC, STAGES, SPLIT>;
    static bool configured = false;
    if (!configured) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN, TPC, STAGES, SPLIT>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    dim3 grid(ksplit, tiles_a * p.n_tiles_b, tap_groups);
    wgrad_kernel<BN, TPC, STAGES, SPLIT><<<grid, 192, L::TOTAL, st>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// tpc = taps sharing one A tile per CTA (3 for a 3x3 kernel row in single-plane mode, else 1)
static int run_wg(WgParams& p, int gh, int gw, int n_img, bool split, cudaStream_t st) {
    p.n_img = n_img;
    const int TW = 1 << p.tw_log2, TH = WG_PIX >> p.tw_log2;
    p.tiles_h = (gh + TH - 1) / TH;
    p.tiles_w = (gw + TW - 1) / TW;
    p.lbo_bytes = g_debug[2] > 0 ? (uint32_t)g_debug[2] : (uint32_t)WG_ATOM_BYTES;
    p.sbo_bytes = g_debug[3] > 0 ? (uint32_t)g_debug[3] : 512u;
    p.layout = g_debug[5] > 0 ? (uint32_t)g_debug[5] : 1u;
    p.scalar_atomics = g_debug[4];
    const int bn = p.cb >= 128 ? 128 : (p.cb >= 64 ? 64 : 32);
    const int tpc = (!split && p.ntaps % 3 == 0) ? 3 : 1;
    p.n_tiles_b = (p.cb + bn - 1) / bn;
    const int tiles_a = (p.ca + 127) / 128;
    const int tap_groups = p.ntaps / tpc;
    const int total_tiles = n_img * p.tiles_h * p.tiles_w;
    const int sms = 148;
    const int blocks_mn = tiles_a * p.n_tiles_b * tap_groups;
    int ksplit = (2 * sms + blocks_mn - 1) / blocks_mn;
    if (ksplit > total_tiles) ksplit = total_tiles;
    if (ksplit < 1) ksplit = 1;
    p.tiles_per_cta = (total_tiles + ksplit - 1) / ksplit;
    ksplit = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    if (split) {
        if (bn == 128) return launch_wg<128, 1, 3, true>(p, ksplit, tiles_a, tap_groups, st);
        if (bn == 64) return launch_wg<64, 1, 4, true>(p, ksplit, tiles_a, tap_groups, st);
        return launch_wg<32, 1, 4, true>(p, ksplit, tiles_a, tap_groups, st);
    }
    if (tpc == 3) {
        if (bn == 128) return launch_wg<128, 3, 3, false>(p, ksplit, tiles_a, tap_groups, st);
        if (bn == 64) return launch_wg<64, 3, 4, false>(p, ksplit, tiles_a, tap_groups, st);
        return launch_wg<32, 3, 4, false>(p, ksplit, tiles_a, tap_groups, st);
    }
    if (bn == 128) return launch_wg<128, 1, 4, false>(p, ksplit, tiles_a, tap_groups, st);
    if (bn == 64) return launch_wg<64, 1, 4, false>(p, ksplit, tiles_a, tap_groups, st);
    return launch_wg<32, 1, 4, false>(p, ksplit, tiles_a, tap_groups, st);
}

// 3xTF32: D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo. Base taps reference base maps [0, nmaps) and weight taps
// [0, ntaps_w); lo activation views live at map + nmaps, lo weights at btap + ntaps_w.
static int expand_split_taps(TgTap* taps, int ntaps, int nmaps, int ntaps_w) {
    for (int i = ntaps - 1; i >= 0; --i) {
        const TgTap t = taps[i];
        TgTap a = t, b = t, c = t;
        b.map = (int16_t)(t.map + nmaps);
        c.btap = t.btap + ntaps_w;
        taps[3 * i] = a;
        taps[3 * i + 1] = b;
        taps[3 * i + 2] = c;
    }
    return 3 * ntaps;
}

// ------------------------------------------------------------------ small re-layout kernels
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int cout, int cin, int kk, int cout_pad,
                                   float* __restrict__ wf, float* __restrict__ wd) {
    const long long total = (long long)kk * cout_pad * cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        const int co = (int)((i / cin) % cout_pad);
        const int tap = (int)(i / ((long long)cin * cout_pad));
        const float v = co < cout ? w[((long long)co * cin + ci) * kk + tap] : 0.f;
        const float hi = tf32_rn(v), lo = v - hi;
        if (wf) {  // [2][tap][co][ci]
            wf[i] = hi;
            wf[total + i] = lo;
        }
        if (wd) {  // [2][tap][ci][co]
            const long long j = ((long long)tap * cin + ci) * cout_pad + co;
            wd[j] = hi;
            wd[total + j] = lo;
        }
    }
}
__global__ void unpack_conv_dw_kernel(const float* __restrict__ dwp, int cout, int cin, int kk, int cout_pad,
                                      float* __restrict__ dw, int accumulate) {
    const long long total = (long long)cout * cin * kk;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(i % kk);
        const int ci = (int)((i / kk) % cin);
        const int co = (int)(i / ((long long)kk * cin));
        const float v = dwp[((long long)tap * cout_pad + co) * cin + ci];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}
__global__ void pack_deconv_w_kernel(const float* __restrict__ w, int cin, int cout, int s, float* __restrict__ wf,
                                     float* __restrict__ wd) {
    const int ss = s * s;
    const long long total = (long long)cin * cout * ss;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ij = (int)(i % ss);
        const int co = (int)((i / ss) % cout);
        const int ci = (int)(i / ((long long)ss * cout));
        const float v = w[i];  // [ci][co][i][j]
        const float hi = tf32_rn(v), lo = v - hi;
        if (wf) {  // [2][(ij, co)][ci]
            const long long j = ((long long)ij * cout + co) * cin + ci;
            wf[j] = hi;
            wf[total + j] = lo;
        }
        if (wd) {  // [2][ij][ci][co]
            const long long j = ((long long)ij * cin + ci) * cout + co;
            wd[j] = hi;
            wd[total + j] = lo;
        }
    }
}
__global__ void unpack_deconv_dw_kernel(const float* __restrict__ dwp, int cin, int cout, int s,
                                        float* __restrict__ dw, int accumulate) {
    const int ss = s * s;
    const long long total = (long long)cin * cout * ss;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ij = (int)(i % ss);
        const int co = (int)((i / ss) % cout);
        const int ci = (int)(i / ((long long)ss * cout));
        const float v = dwp[((long long)ij * cin + ci) * cout + co];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, float* w_fwd, float* w_dgrad,
                         a2x_stream_t stream) {
    A2X_REQUIRE(w_oihw && cout > 0 && cin > 0 && (ksize == 1 || ksize == 3) && cout_pad >= cout, "bad pack args");
    const int kk = ksize * ksize;
    pack_conv_w_kernel<<<grid_for((long long)kk * cout_pad * cin), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, cout, cin, kk, cout_pad, w_fwd, w_dgrad);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_unpack_conv_wgrad(const float* dw_packed, int cout, int cin, int ksize, int cout_pad, float* dw_oihw,
                          int accumulate, a2x_stream_t stream) {
    A2X_REQUIRE(dw_packed && dw_oihw && cout > 0 && cin > 0 && cout_pad >= cout, "bad unpack args");
    const int kk = ksize * ksize;
    unpack_conv_dw_kernel<<<grid_for((long long)kk * cout * cin), 256, 0, (cudaStream_t)stream>>>(
        dw_packed, cout, cin, kk, cout_pad, dw_oihw, accumulate);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_pack_deconv_weight(const float* w_iohw, int cin, int cout, int s, float* w_fwd, float* w_dgrad,
                           a2x_stream_t stream) {
    A2X_REQUIRE(w_iohw && cin > 0 && cout > 0 && s > 0, "bad pack args");
    pack_deconv_w_kernel<<<grid_for((long long)cin * cout * s * s), 256, 0, (cudaStream_t)stream>>>(w_iohw, cin, cout,
                                                                                                  s, w_fwd, w_dgrad);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_unpack_deconv_wgrad(const float* dw_packed, int cin, int cout, int s, float* dw_iohw, int accumulate,
                            a2x_stream_t stream) {
    A2X_REQUIRE(dw_packed && dw_iohw && cin > 0 && cout > 0 && s > 0, "bad unpack args");
    unpack_deconv_dw_kernel<<<grid_for((long long)cin * cout * s * s), 256, 0, (cudaStream_t)stream>>>(
        dw_packed, cin, cout, s, dw_iohw, accumulate);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_conv2d_fwd(const a2x_conv_shape* s, const float* x, const float* x_lo, int x_cs, const float* w_fwd, float* y,
                   float* y_lo, int y_cs, const float* scale, const float* shift, int relu, double* stats,
                   a2x_stream_t stream) {
    if (int r = check_shape(s, false)) return r;
    A2X_REQUIRE(x && w_fwd && y && x_cs >= s->cin && y_cs >= s->cout && x_cs % 4 == 0 && y_cs % 4 == 0,
                "bad conv2d_fwd pointers/strides");
    const int ho = (s->h - 1) / s->stride + 1, wo = (s->w - 1) / s->stride + 1;
    const int kk = s->ksize * s->ksize;
    TgParams p{};
    p.tw_log2 = pick_tw_log2(ho, wo, TG_BM, 4, 7);
    const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
    const int nmaps = s->stride == 1 ? 1 : 4;
    if (int r = build_fwd_maps(s, x, x_cs, p.amap, TW, TH)) return r;
    p.ntaps = build_fwd_taps(s, p.taps);
    if (x_lo) {
        if (int r = build_fwd_maps(s, x_lo, x_cs, p.amap + nmaps, TW, TH)) return r;
        p.ntaps = expand_split_taps(p.taps, p.ntaps, nmaps, kk);
    }
    p.kchunks = s->cin / 32;
    if (int r = make_w_map(&p.bmap, w_fwd, 2 * kk, s->cout, s->cin, bn_for(s->cout))) return r;
    set_plain_out(p, y, ho, wo, y_cs, 1, 0, 0);
    p.out_lo = y_lo;
    p.scale = scale;
    p.shift = shift;
    p.relu = relu;
    p.accumulate = 0;
    p.stats = stats;
    p.stat_c = s->cout;
    A2X_REQUIRE(!stats || (!scale && !shift && !relu), "conv2d_fwd: fused statistics are of the raw conv output");
    return run_tg(p, ho, wo, s->n, s->cout, (cudaStream_t)stream);
}

int a2x_conv2d_dgrad(const a2x_conv_shape* s, const float* dy, const float* dy_lo, int dy_cs, const float* w_dgrad,
                     float* dx, int dx_cs, int accumulate, a2x_stream_t stream) {
    if (int r = check_shape(s, false)) return r;
    A2X_REQUIRE(dy && w_dgrad && dx && dy_cs >= s->cout && dx_cs >= s->cin && dy_cs % 4 == 0 && dx_cs % 4 == 0,
                "bad conv2d_dgrad pointers/strides");
    const int ho = (s->h - 1) / s->stride + 1, wo = (s->w - 1) / s->stride + 1;
    const int kk = s->ksize * s->ksize;
    const int n_class = s->stride == 1 ? 1 : 4;
    for (int cls = 0; cls < n_class; ++cls) {
        const int hp = cls >> 1, wp = cls & 1;
        const int step = s->stride;
        const int gh = (s->h - hp + step - 1) / step, gw = (s->w - wp + step - 1) / step;
        if (gh <= 0 || gw <= 0) continue;
        TgParams p{};
        p.tw_log2 = pick_tw_log2(gh, gw, TG_BM, 4, 7);
        const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
        if (int r = make_act_map(&p.amap[0], dy, s->n, ho, wo, s->cout, dy_cs, 1, 0, 0, TW, TH)) return r;
        p.ntaps = 0;
        for (int r = 0; r < s->ksize; ++r) {
            if (step == 2 && ((r + 1) & 1) != hp) continue;  // (r - 1) parity must equal hp
            for (int c = 0; c < s->ksize; ++c) {
                if (step == 2 && ((c + 1) & 1) != wp) continue;
                TgTap t{};
                t.map = 0;
                if (step == 1) {
                    t.dh = (int16_t)(s->ksize == 3 ? 1 - r : 0);
                    t.dw = (int16_t)(s->ksize == 3 ? 1 - c : 0);
                } else {
                    t.dh = (int16_t)((hp == 1 && r == 0) ? 1 : 0);
                    t.dw = (int16_t)((wp == 1 && c == 0) ? 1 : 0);
                }
                t.btap = r * s->ksize + c;
                p.taps[p.ntaps++] = t;
            }
        }
        if (dy_lo) {
            if (int r = make_act_map(&p.amap[1], dy_lo, s->n, ho, wo, s->cout, dy_cs, 1, 0, 0, TW, TH)) return r;
            p.ntaps = expand_split_taps(p.taps, p.ntaps, 1, kk);
        }
        p.kchunks = s->cout / 32;
        if (int r = make_w_map(&p.bmap, w_dgrad, 2 * kk, s->cin, s->cout, bn_for(s->cin))) return r;
        set_plain_out(p, dx, s->h, s->w, dx_cs, step, hp, wp);
        p.accumulate = accumulate;
        if (int r = run_tg(p, gh, gw, s->n, s->cin, (cudaStream_t)stream)) return r;
    }
    return 0;
}

int a2x_conv2d_wgrad(const a2x_conv_shape* s, const float* x, const float* x_lo, int x_cs, const float* dy,
                     const float* dy_lo, int dy_cs, float* dw_packed, a2x_stream_t stream) {
    if (int r = check_shape(s, false)) return r;
    A2X_REQUIRE(x && dy && dw_packed && x_cs >= s->cin && dy_cs >= s->cout && x_cs % 4 == 0 && dy_cs % 4 == 0,
                "bad conv2d_wgrad pointers/strides");
    A2X_REQUIRE((x_lo == nullptr) == (dy_lo == nullptr), "wgrad needs both or neither lo planes");
    const int ho = (s->h - 1) / s->stride + 1, wo = (s->w - 1) / s->stride + 1;
    WgParams p{};
    p.tw_log2 = pick_tw_log2(ho, wo, WG_PIX, 2, 5);
    const int TW = 1 << p.tw_log2, TH = WG_PIX >> p.tw_log2;
    p.nmaps_b = s->stride == 1 ? 1 : 4;
    if (int r = make_act_map(&p.amap[0], dy, s->n, ho, wo, s->cout, dy_cs, 1, 0, 0, TW, TH, 1)) return r;
    if (int r = build_fwd_maps(s, x, x_cs, p.bmap, TW, TH, 1)) return r;
    if (x_lo) {
        if (int r = make_act_map(&p.amap[1], dy_lo, s->n, ho, wo, s->cout, dy_cs, 1, 0, 0, TW, TH, 1)) return r;
        if (int r = build_fwd_maps(s, x_lo, x_cs, p.bmap + p.nmaps_b, TW, TH, 1)) return r;
    }
    p.ntaps = build_fwd_taps(s, p.taps);
    p.ca = s->cout;
    p.cb = s->cin;
    p.dw = dw_packed;
    return run_wg(p, ho, wo, s->n, x_lo != nullptr, (cudaStream_t)stream);
}

int a2x_deconv_fwd(const a2x_conv_shape* s, const float* x, const float* x_lo, int x_cs, const float* w_fwd, float* y,
                   float* y_lo, int y_cs, const float* scale, const float* shift, int relu, double* stats,
                   a2x_stream_t stream) {
    if (int r = check_shape(s, true)) return r;
    A2X_REQUIRE(x && w_fwd && y && x_cs >= s->cin && y_cs >= s->cout && x_cs % 4 == 0 && y_cs % 4 == 0,
                "bad deconv_fwd pointers/strides");
    const int st = s->stride;
    TgParams p{};
    p.tw_log2 = pick_tw_log2(s->h, s->w, TG_BM, 4, 7);
    const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
    if (int r = make_act_map(&p.amap[0], x, s->n, s->h, s->w, s->cin, x_cs, 1, 0, 0, TW, TH)) return r;
    p.ntaps = 1;
    p.taps[0] = TgTap{0, 0, 0, 0, 0, 0};
    if (x_lo) {
        if (int r = make_act_map(&p.amap[1], x_lo, s->n, s->h, s->w, s->cin, x_cs, 1, 0, 0, TW, TH)) return r;
        p.ntaps = expand_split_taps(p.taps, 1, 1, 1);
    }
    p.kchunks = s->cin / 32;
    const int ncols = st * st * s->cout;
    if (int r = make_w_map(&p.bmap, w_fwd, 2, ncols, s->cin, bn_for(ncols))) return r;
    const long long W2 = (long long)s->w * st;
    p.out = y;
    p.out_lo = y_lo;
    p.osn = (long long)s->h * st * W2 * y_cs;
    p.osh = (long long)st * W2 * y_cs;
    p.osw = (long long)st * y_cs;
    p.sub_c = s->cout;
    p.sub_s = st;
    p.sub_sh = W2 * y_cs;
    p.sub_sw = y_cs;
    p.scale = scale;
    p.shift = shift;
    p.relu = relu;
    p.accumulate = 0;
    p.stats = stats;
    p.stat_c = s->cout;
    A2X_REQUIRE(!stats || (!scale && !shift && !relu), "deconv_fwd: fused statistics are of the raw output");
    return run_tg(p, s->h, s->w, s->n, ncols, (cudaStream_t)stream);
}

int a2x_deconv_dgrad(const a2x_conv_shape* s, const float* dy, const float* dy_lo, int dy_cs, const float* w_dgrad,
                     float* dx, int dx_cs, int accumulate, a2x_stream_t stream) {
    if (int r = check_shape(s, true)) return r;
    A2X_REQUIRE(dy && w_dgrad && dx && dy_cs >= s->cout && dx_cs >= s->cin && dy_cs % 4 == 0 && dx_cs % 4 == 0,
                "bad deconv_dgrad pointers/strides");
    const int st = s->stride;
    // one launch per sub-row i; its s sub-columns j are the taps, each through its own strided view of dy
    for (int i = 0; i < st; ++i) {
        TgParams p{};
        p.tw_log2 = pick_tw_log2(s->h, s->w, TG_BM, 4, 7);
        const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
        p.ntaps = 0;
        for (int j = 0; j < st; ++j) {
            if (int r = make_act_map(&p.amap[j], dy, s->n, s->h * st, s->w * st, s->cout, dy_cs, st, i, j, TW, TH))
                return r;
            if (dy_lo)
                if (int r = make_act_map(&p.amap[st + j], dy_lo, s->n, s->h * st, s->w * st, s->cout, dy_cs, st, i, j,
                                         TW, TH))
                    return r;
            TgTap t{};
            t.map = (int16_t)j;
            t.btap = i * st + j;
            p.taps[p.ntaps++] = t;
        }
        if (dy_lo) p.ntaps = expand_split_taps(p.taps, p.ntaps, st, st * st);
        p.kchunks = s->cout / 32;
        if (int r = make_w_map(&p.bmap, w_dgrad, 2 * st * st, s->cin, s->cout, bn_for(s->cin))) return r;
        set_plain_out(p, dx, s->h, s->w, dx_cs, 1, 0, 0);
        p.accumulate = (i > 0) ? 1 : accumulate;
        if (int r = run_tg(p, s->h, s->w, s->n, s->cin, (cudaStream_t)stream)) return r;
    }
    return 0;
}

int a2x_deconv_wgrad(const a2x_conv_shape* s, const float* x, const float* x_lo, int x_cs, const float* dy,
                     const float* dy_lo, int dy_cs, float* dw_packed, a2x_stream_t stream) {
    if (int r = check_shape(s, true)) return r;
    A2X_REQUIRE(x && dy && dw_packed && x_cs >= s->cin && dy_cs >= s->cout && x_cs % 4 == 0 && dy_cs % 4 == 0,
                "bad deconv_wgrad pointers/strides");
    A2X_REQUIRE((x_lo == nullptr) == (dy_lo == nullptr), "wgrad needs both or neither lo planes");
    const int st = s->stride;
    for (int i = 0; i < st; ++i) {
        WgParams p{};
        p.tw_log2 = pick_tw_log2(s->h, s->w, WG_PIX, 2, 5);
        const int TW = 1 << p.tw_log2, TH = WG_PIX >> p.tw_log2;
        p.nmaps_b = st;
        if (int r = make_act_map(&p.amap[0], x, s->n, s->h, s->w, s->cin, x_cs, 1, 0, 0, TW, TH, 1)) return r;
        if (x_lo)
            if (int r = make_act_map(&p.amap[1], x_lo, s->n, s->h, s->w, s->cin, x_cs, 1, 0, 0, TW, TH, 1)) return r;
        p.ntaps = 0;
        for (int j = 0; j < st; ++j) {
            if (int r = make_act_map(&p.bmap[j], dy, s->n, s->h * st, s->w * st, s->cout, dy_cs, st, i, j, TW, TH, 1))
                return r;
            if (dy_lo)
                if (int r = make_act_map(&p.bmap[st + j], dy_lo, s->n, s->h * st, s->w * st, s->cout, dy_cs, st, i, j,
                                         TW, TH, 1))
                    return r;
            TgTap t{};
            t.map = (int16_t)j;
            p.taps[p.ntaps++] = t;
        }
        p.ca = s->cin;
        p.cb = s->cout;
        p.dw = dw_packed + (long long)i * st * s->cin * s->cout;  // [(i, j)][ci][co]
        if (int r = run_wg(p, s->h, s->w, s->n, x_lo != nullptr, (cudaStream_t)stream)) return r;
    }
    return 0;
}

}  // extern "C"
