// C-ABI launchers for the dense contractions (conv / transposed conv: forward, data gradient, weight gradient).
// All of them lower onto the two tcgen05 kernels in tapgemm.cuh / wgrad.cuh by choosing tensor-map views and taps.
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "tapgemm.cuh"
#include "tapgemm_halo.cuh"
#include "wgrad.cuh"

namespace a2x {

extern int g_debug[16];

// ------------------------------------------------------------------ tensor-map views of an NHWC tensor
// Rank-5 view (c, w, x, h, n) of every `step`-th pixel starting at (h_off, w_off).
// fp32: box = (32, box_w, 1, box_h, 1); bf16: box = (64, ...). `cs` = pixel stride in elements.
static int make_act_map(CUtensorMap* m, const void* base, int n, int h, int w, int c, int cs, int step, int h_off,
                        int w_off, int box_w, int box_h, int atom32 = 0, int bf16 = 0) {
    const int hp = (h - h_off + step - 1) / step;
    const int wp = (w - w_off + step - 1) / step;
    if (hp <= 0 || wp <= 0) {
        set_error("empty activation view");
        return 1;
    }
    const size_t es = bf16 ? 2 : 4;
    const char* b = (const char*)base + ((long long)h_off * w + w_off) * cs * es;
    uint64_t dims[5] = {(uint64_t)c, (uint64_t)wp, 1, (uint64_t)hp, (uint64_t)n};
    uint64_t str[4] = {(uint64_t)step * cs * es, (uint64_t)step * cs * es, (uint64_t)step * w * cs * es,
                       (uint64_t)h * w * cs * es};
    uint32_t box[5] = {(uint32_t)(bf16 ? 64 : 32), (uint32_t)box_w, 1, (uint32_t)box_h, 1};
    return encode_tmap_f32(m, b, 5, dims, str, box, atom32, bf16);
}

// weights [taps][rows][k] (k contiguous); box = (32 | 64, bn, 1)
static int make_w_map(CUtensorMap* m, const void* base, int taps, int rows, int k, int bn, int bf16 = 0) {
    const size_t es = bf16 ? 2 : 4;
    uint64_t dims[3] = {(uint64_t)k, (uint64_t)rows, (uint64_t)taps};
    uint64_t str[2] = {(uint64_t)k * es, (uint64_t)rows * k * es};
    uint32_t box[3] = {(uint32_t)(bf16 ? 64 : 32), (uint32_t)(bn < rows ? bn : rows), 1};
    return encode_tmap_f32(m, base, 3, dims, str, box, 0, bf16);
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
    int n_tiles = p.n_img * p.tiles_h * p.tiles_w * n_col_tiles;
    const int cap = g_debug[8] > 0 ? g_debug[8] : 148;  // persistent: one CTA per SM
    dim3 grid(n_tiles < cap ? n_tiles : cap, 1, 1);
    tapgemm_kernel<BN, STAGES><<<grid, 64 + TG_EPW * 32, L::TOTAL, st>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int bn_for(int ncols) {
    int bn = ncols >= 256 ? 256 : ncols >= 128 ? 128 : ncols >= 64 ? 64 : 32;
    if (g_debug[1] > 0) bn = g_debug[1];
    return bn;
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
    const int bn = bn_for(ncols);
    if (ncols % 32 != 0 || bn > ncols) {
        set_error("column count %d not a multiple of 32 (tile %d)", ncols, bn);
        return 1;
    }
    const int tiles_n = (ncols + bn - 1) / bn;
    switch (bn) {
        case 256: return launch_tg<256, 4>(p, tiles_n, st);
        case 128: return launch_tg<128, 6>(p, tiles_n, st);
        case 64: return launch_tg<64, 8>(p, tiles_n, st);
        case 32: return launch_tg<32, 8>(p, tiles_n, st);
    }
    set_error("bad bn %d", bn);
    return 1;
}

static void set_plain_out(TgParams& p, const a2x_output* y, int h, int w, int step, int h_off, int w_off) {
    const long long o = ((long long)h_off * w + w_off) * y->cs;
    p.out.hi = y->hi ? y->hi + o : nullptr;   // null fp32 plane: only the split planes are stored (store_split4)
    p.out.b16 = y->b16 ? (__nv_bfloat16*)y->b16 + o : nullptr;
    p.out.ps = y->b16_plane;
    p.osn = (long long)h * w * y->cs;
    p.osh = (long long)step * w * y->cs;
    p.osw = (long long)step * y->cs;
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
        if (!((s->ksize == 1 && s->stride == 1) || (s->ksize == 3 && (s->stride == 1 || s->stride == 2)) ||
              (s->ksize == 7 && s->stride == 2))) {
            set_error("conv needs (k=1,s=1), (k=3,s in {1,2}) or the forward-only stem (k=7,s=2), got k=%d s=%d", s->ksize, s->stride);
            return 1;
        }
    }
    return 0;
}

static int check_operand(const a2x_operand* x, int c, const char* what) {
    if (!x || !x->hi || x->cs < c || x->cs % 4) {
        set_error("%s: bad operand (null / pixel stride < channels / stride not a multiple of 4)", what);
        return 1;
    }
    if (x->b16 && (c % 64 || x->cs % 8)) {
        set_error("%s: split operands need channel counts that are multiples of 64 (got %d)", what, c);
        return 1;
    }
    return 0;
}

template <int BN, int NB>
static int launch_th(const ThParams& p, int n_col_tiles, cudaStream_t st) {
    using L = ThSmem<BN, NB>;
    static bool configured = false;
    if (!configured) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_halo_kernel<BN, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            L::TOTAL));
        configured = true;
    }
    int n_tiles = p.n_img * p.tiles_h * p.tiles_w * n_col_tiles;
    const int cap = g_debug[8] > 0 ? g_debug[8] : 148;
    dim3 grid(n_tiles < cap ? n_tiles : cap, 1, 1);
    tapgemm_halo_kernel<BN, NB><<<grid, 192, L::TOTAL, st>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// halo-box view (c, w, x, h, n) of a dense (step 1) NHWC tensor: box = (32|64, 10, 1, 18, 1)
static int make_halo_map(CUtensorMap* m, const void* base, int n, int h, int w, int c, int cs, int bf16) {
    const size_t es = bf16 ? 2 : 4;
    uint64_t dims[5] = {(uint64_t)c, (uint64_t)w, 1, (uint64_t)h, (uint64_t)n};
    uint64_t str[4] = {(uint64_t)cs * es, (uint64_t)cs * es, (uint64_t)w * cs * es, (uint64_t)h * w * cs * es};
    uint32_t box[5] = {(uint32_t)(bf16 ? 64 : 32), TH_COLS, 1, TH_ROWS, 1};
    return encode_tmap_f32(m, base, 5, dims, str, box, 0, bf16);
}

// 3x3 stride-1 window GEMM through the halo kernel. sign = +1: forward taps (r-1, c-1); -1: data-gradient (1-r, 1-c).
static int run_halo(const a2x_operand* a, int n, int h, int w, int k_ch, int ncols, const a2x_weights* wts, int sign,
                    const a2x_output* y, const float* scale, const float* shift, int relu, int accumulate,
                    double* stats, cudaStream_t st, const a2x_bn_bwd_stats* bst = nullptr) {
    ThParams p{};
    const int bn = bn_for(ncols);
    if (a->b16) {
        // two A passes: h16 against the 18 weight taps (h16 then l16 planes), l16 against the 9 h16 taps
        const __nv_bfloat16* b = (const __nv_bfloat16*)a->b16;
        if (int r = make_halo_map(&p.amap[0], b, n, h, w, k_ch, a->cs, 1)) return r;                // h16
        if (int r = make_halo_map(&p.amap[1], b + a->b16_plane, n, h, w, k_ch, a->cs, 1)) return r;  // l16
        if (int r = make_w_map(&p.bmap, wts->w16, 18, ncols, k_ch, bn, 1)) return r;
        p.npass = 2;
        p.pass_taps[0] = 18;
        p.pass_taps[1] = 9;
        p.kind = 1;
        p.kchunks = k_ch / 64;
    } else {
        if (int r = make_halo_map(&p.amap[0], a->hi, n, h, w, k_ch, a->cs, 0)) return r;
        if (int r = make_w_map(&p.bmap, wts->w32, 9, ncols, k_ch, bn)) return r;
        p.npass = 1;
        p.pass_taps[0] = 9;
        p.kind = 0;
        p.kchunks = k_ch / 32;
    }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            p.dh[r * 3 + c] = (int8_t)(sign * (r - 1));
            p.dw[r * 3 + c] = (int8_t)(sign * (c - 1));
        }
    p.n_img = n;
    p.gh = h;
    p.gw = w;
    p.tiles_h = (h + 15) / 16;
    p.tiles_w = (w + 7) / 8;
    p.out.hi = y->hi;
    p.out.b16 = (__nv_bfloat16*)y->b16;
    p.out.ps = y->b16_plane;
    p.osn = (long long)h * w * y->cs;
    p.osh = (long long)w * y->cs;
    p.osw = y->cs;
    p.ncols = ncols;
    p.scale = scale;
    p.shift = shift;
    p.relu = relu;
    p.accumulate = accumulate;
    p.stats = stats;
    p.stat_c = ncols;
    if (bst != nullptr) {  // fused BN+ReLU backward reduction of the consumer layer (data-gradient epilogue)
        if (!bst->z || !bst->scale || !bst->shift || !bst->mean || !bst->invstd || !bst->sums || bst->z_cs != y->cs) {
            set_error("conv2d_dgrad: bad bn_bwd_stats (z must share the output's pixel stride)");
            return 1;
        }
        p.stats = bst->sums;
        p.bz = bst->z; p.bscale = bst->scale; p.bshift = bst->shift; p.bmean = bst->mean; p.binvstd = bst->invstd;
    }
    p.base_offset_mode = g_debug[6];
    if (g_debug[9]) p.relu = 77;
    const int tiles_n = (ncols + bn - 1) / bn;
    switch (bn) {
        case 256: return launch_th<256, 5>(p, tiles_n, st);
        case 128: return launch_th<128, 10>(p, tiles_n, st);
        case 64: return launch_th<64, 18>(p, tiles_n, st);
        case 32: return launch_th<32, 18>(p, tiles_n, st);
    }
    set_error("bad bn %d", bn);
    return 1;
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

// maps of the input x for a forward conv: [0, nmaps) fp32 hi; in split mode [nmaps, 2 nmaps) h16, [2 nmaps, 3 nmaps) l16
static int build_fwd_maps(const a2x_conv_shape* s, const a2x_operand* x, CUtensorMap* maps, int box_w, int box_h,
                          int atom32 = 0, CUtensorMap* maps_h16 = nullptr, CUtensorMap* maps_l16 = nullptr) {
    const int nm = s->stride == 1 ? 1 : 4;
    for (int i = 0; i < nm; ++i) {
        const int step = s->stride, hp = step == 1 ? 0 : i >> 1, wp = step == 1 ? 0 : i & 1;
        if (hp >= s->h || wp >= s->w) continue;
        if (int r = make_act_map(&maps[i], x->hi, s->n, s->h, s->w, s->cin, x->cs, step, hp, wp, box_w, box_h, atom32))
            return r;
        if (x->b16 && maps_h16) {
            const __nv_bfloat16* b = (const __nv_bfloat16*)x->b16;
            if (int r = make_act_map(&maps_h16[i], b, s->n, s->h, s->w, s->cin, x->cs, step, hp, wp, box_w, box_h, 0, 1))
                return r;
            if (int r = make_act_map(&maps_l16[i], b + x->b16_plane, s->n, s->h, s->w, s->cin, x->cs, step, hp, wp,
                                     box_w, box_h, 0, 1))
                return r;
        }
    }
    return 0;
}

// split mode: D = A_h*B_h + A_l*B_h + A_h*B_l, all kind::f16 (bf16). Base taps reference fp32 maps [0, nmaps) and
// weight taps [0, ntaps_w); h16 views live at map + nmaps, l16 views at map + 2 nmaps; the bf16 weight tensor carries
// the h16 taps first, then the l16 taps (+ ntaps_w).
static int expand_split_taps(TgTap* taps, int ntaps, int nmaps, int ntaps_w) {
    for (int i = ntaps - 1; i >= 0; --i) {
        const TgTap t = taps[i];
        TgTap a = t, b = t, c = t;
        a.kind = b.kind = c.kind = 1;
        a.map = (int16_t)(t.map + nmaps);      // A h16 x W h16
        b.map = (int16_t)(t.map + 2 * nmaps);  // A l16 x W h16
        c.map = (int16_t)(t.map + nmaps);      // A h16 x W l16
        c.btap = t.btap + ntaps_w;
        taps[3 * i] = a;
        taps[3 * i + 1] = b;
        taps[3 * i + 2] = c;
    }
    return 3 * ntaps;
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

static int run_wg(WgParams& p, int gh, int gw, int n_img, bool split, cudaStream_t st) {
    p.n_img = n_img;
    const int TW = 1 << p.tw_log2, TH = WG_PIX >> p.tw_log2;
    p.tiles_h = (gh + TH - 1) / TH;
    p.tiles_w = (gw + TW - 1) / TW;
    p.lbo_bytes = g_debug[2] > 0 ? (uint32_t)g_debug[2] : (uint32_t)WG_ATOM_BYTES;
    p.sbo_bytes = g_debug[3] > 0 ? (uint32_t)g_debug[3] : 512u;
    p.layout = g_debug[5] > 0 ? (uint32_t)g_debug[5] : 1u;
    // one tap (1x1 conv = the token-wise linears), split mode, wide b side: 256-column tiles halve the re-reads of the A
    // operand (these launches are bound by operand traffic: 761 MB of DRAM reads for 504 MB of operands at 128 x 128)
    const bool wide = split && p.ntaps == 1 && p.cb % 256 == 0 && g_debug[4] != 1;
    const int bn = wide ? 256 : p.cb >= 128 ? 128 : (p.cb >= 64 ? 64 : 32);
    const int tpc = (p.ntaps % 3 == 0 && bn <= 128) ? 3 : 1;
    p.n_tiles_b = (p.cb + bn - 1) / bn;
    const int tiles_a = (p.ca + 127) / 128;
    const int tap_groups = p.ntaps / tpc;
    const int total_tiles = n_img * p.tiles_h * p.tiles_w;
    const int sms = 148;
    const int blocks_mn = tiles_a * p.n_tiles_b * tap_groups;
    int ksplit = (tpc * bn >= 256) ? sms / blocks_mn : (2 * sms + blocks_mn - 1) / blocks_mn;  // 512 TMEM columns / 192 KB of stages => 1 CTA / SM
    if (ksplit > total_tiles) ksplit = total_tiles;
    if (ksplit < 1) ksplit = 1;
    p.tiles_per_cta = (total_tiles + ksplit - 1) / ksplit;
    ksplit = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    if (split) {
        if (tpc == 3) {
            if (bn == 128) return launch_wg<128, 3, 3, true>(p, ksplit, tiles_a, tap_groups, st);
            if (bn == 64) return launch_wg<64, 3, 4, true>(p, ksplit, tiles_a, tap_groups, st);
        } else {
            if (bn == 256) return launch_wg<256, 1, 4, true>(p, ksplit, tiles_a, tap_groups, st);
            if (bn == 128) return launch_wg<128, 1, 4, true>(p, ksplit, tiles_a, tap_groups, st);
            if (bn == 64) return launch_wg<64, 1, 4, true>(p, ksplit, tiles_a, tap_groups, st);
        }
        set_error("split wgrad needs cb >= 64");
        return 1;
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

template <int BN, int STAGES>
static int launch_wr(const WrParams& p, int ksplit, int blocks_mn, cudaStream_t st) {
    using L = WrSmem<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(wgrad_row_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            L::TOTAL));
        configured = true;
    }
    dim3 grid(ksplit, blocks_mn, 3);
    wgrad_row_kernel<BN, STAGES><<<grid, 192, L::TOTAL, st>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// 3x3 stride-1 weight gradient in split mode through the row-halo kernel
static int run_wr(const a2x_conv_shape* s, const a2x_operand* x, const a2x_operand* dy, float* dw_packed, cudaStream_t st) {
    WrParams p{};
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x->b16;
    const __nv_bfloat16* db = (const __nv_bfloat16*)dy->b16;
    for (int pl = 0; pl < 2; ++pl) {
        if (int r = make_act_map(&p.amap16[pl], db + pl * dy->b16_plane, s->n, s->h, s->w, s->cout, dy->cs, 1, 0, 0,
                                 WG_PIX, 1, 0, 1))
            return r;
        if (int r = make_act_map(&p.bmap16[pl], xb + pl * x->b16_plane, s->n, s->h, s->w, s->cin, x->cs, 1, 0, 0,
                                 WR_B_ROWS, 1, 0, 1))
            return r;
    }
    p.ca = s->cout;
    p.cb = s->cin;
    p.n_img = s->n;
    p.gh = s->h;
    p.tiles_w = (s->w + WG_PIX - 1) / WG_PIX;
    p.dw = dw_packed;
    const int bn = p.cb >= 128 ? 128 : 64;
    p.n_tiles_b = (p.cb + bn - 1) / bn;
    const int tiles_a = (p.ca + 127) / 128;
    const int blocks_mn = tiles_a * p.n_tiles_b;
    const int total_tiles = p.n_img * p.gh * p.tiles_w;
    // BN = 128 allocates all 512 TMEM columns (one CTA per SM): a single wave of <= 148 CTAs; BN = 64 fits two per SM
    int ksplit = (bn == 128 ? 148 : 296) / (blocks_mn * 3);
    if (g_debug[10] > 0) ksplit = g_debug[10];
    if (ksplit > total_tiles) ksplit = total_tiles;
    if (ksplit < 1) ksplit = 1;
    p.tiles_per_cta = (total_tiles + ksplit - 1) / ksplit;
    ksplit = (total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    if (bn == 128) return launch_wr<128, 5>(p, ksplit, blocks_mn, st);
    return launch_wr<64, 4>(p, ksplit, blocks_mn, st);
}

// ------------------------------------------------------------------ small re-layout kernels
__device__ __forceinline__ void put_w(float* w32, __nv_bfloat16* w16, long long plane, long long j, float v) {
    if (w32) w32[j] = tf32_rn(v);  // single-plane (kind::tf32) operand: pre-rounded, the MMA would truncate
    if (w16) split_bf16(v, w16[j], w16[plane + j]);
}
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int cout, int cin, int kk, int cout_pad,
                                   float* __restrict__ wf, __nv_bfloat16* __restrict__ wf16, float* __restrict__ wd,
                                   __nv_bfloat16* __restrict__ wd16) {
    const long long total = (long long)kk * cout_pad * cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        const int co = (int)((i / cin) % cout_pad);
        const int tap = (int)(i / ((long long)cin * cout_pad));
        const float v = co < cout ? w[((long long)co * cin + ci) * kk + tap] : 0.f;
        put_w(wf, wf16, total, i, v);                                                  // [tap][co][ci]
        put_w(wd, wd16, total, ((long long)tap * cin + ci) * cout_pad + co, v);        // [tap][ci][co]
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
                                     __nv_bfloat16* __restrict__ wf16, float* __restrict__ wd,
                                     __nv_bfloat16* __restrict__ wd16) {
    const int ss = s * s;
    const long long total = (long long)cin * cout * ss;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ij = (int)(i % ss);
        const int co = (int)((i / ss) % cout);
        const int ci = (int)(i / ((long long)ss * cout));
        const float v = w[i];                                                          // [ci][co][i][j]
        put_w(wf, wf16, total, ((long long)ij * cout + co) * cin + ci, v);             // [(ij, co)][ci]
        put_w(wd, wd16, total, ((long long)ij * cin + ci) * cout + co, v);             // [ij][ci][co]
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

// ------------------------------------------------------------------ batched re-layout (one launch per step)
// The per-layer pack / unpack kernels above move a few hundred KB each: launch latency, not bandwidth, is what they
// cost inside the step. These two kernels walk a device-resident job table instead (<= 128 jobs, staged in smem).
__global__ void __launch_bounds__(256) pack_batched_kernel(const a2x_pack_job* __restrict__ jobs, int njobs) {
    // Every weight is written in TWO layouts (forward operand [tap][co][ci], data-gradient operand [tap][ci][co]); each
    // layout gets its own sweep with the DESTINATION index fastest over the threads, so all stores coalesce (the strided
    // reads of the 29 MB of parameters are served by L2). Null f32 / d32 pointers (split mode: only the bf16 planes feed
    // the GEMMs) skip those planes.
    __shared__ a2x_pack_job sj[128];
    for (int i = threadIdx.x; i < njobs; i += blockDim.x) sj[i] = jobs[i];
    __syncthreads();
    const long long total = sj[njobs - 1].elem_begin + sj[njobs - 1].elems;
    for (int sweep = 0; sweep < 2; ++sweep) {
        int j = 0;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
             i += (long long)gridDim.x * blockDim.x) {
            while (i >= sj[j].elem_begin + sj[j].elems) ++j;  // i grows monotonically per thread
            const a2x_pack_job& J = sj[j];
            const long long e = i - J.elem_begin;
            if (J.kind == 2) {  // plain vector copy (fused head bias)
                if (sweep == 0) J.f32[J.row0 + e] = J.src[e];
                continue;
            }
            if (J.kind == 0) {  // conv OIHW [rows][cin][kk] -> [tap][row0 + co][ci] and [tap][ci][row0 + co]
                const int cin = J.b, rows = J.a, kk = J.kk;
                const long long plane = (long long)kk * J.cout_pad * cin;
                if (sweep == 0) {
                    const int ci = (int)(e % cin);
                    const int co = (int)((e / cin) % rows);
                    const int tap = (int)(e / ((long long)cin * rows));
                    const float w = J.src[((long long)co * cin + ci) * kk + tap];
                    put_w(J.f32, (__nv_bfloat16*)J.f16, plane, ((long long)tap * J.cout_pad + J.row0 + co) * cin + ci, w);
                } else {
                    const int co = (int)(e % rows);
                    const int ci = (int)((e / rows) % cin);
                    const int tap = (int)(e / ((long long)cin * rows));
                    const float w = J.src[((long long)co * cin + ci) * kk + tap];
                    put_w(J.d32, (__nv_bfloat16*)J.d16, plane, ((long long)tap * cin + ci) * J.cout_pad + J.row0 + co, w);
                }
            } else {  // deconv [ci][co][i][j] -> [(ij, co)][ci] and [ij][ci][co]
                const int cin = J.a, cout = J.b, ss = J.kk;
                const long long plane = (long long)cin * cout * ss;
                if (sweep == 0) {
                    const int ci = (int)(e % cin);
                    const int co = (int)((e / cin) % cout);
                    const int ij = (int)(e / ((long long)cin * cout));
                    const float w = J.src[((long long)ci * cout + co) * ss + ij];
                    put_w(J.f32, (__nv_bfloat16*)J.f16, plane, ((long long)ij * cout + co) * cin + ci, w);
                } else {
                    const int co = (int)(e % cout);
                    const int ci = (int)((e / cout) % cin);
                    const int ij = (int)(e / ((long long)cin * cout));
                    const float w = J.src[((long long)ci * cout + co) * ss + ij];
                    put_w(J.d32, (__nv_bfloat16*)J.d16, plane, ((long long)ij * cin + ci) * cout + co, w);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) unpack_batched_kernel(const a2x_unpack_job* __restrict__ jobs, int njobs) {
    __shared__ a2x_unpack_job sj[128];
    for (int i = threadIdx.x; i < njobs; i += blockDim.x) sj[i] = jobs[i];
    __syncthreads();
    const long long total = sj[njobs - 1].elem_begin + sj[njobs - 1].elems;
    int j = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        while (i >= sj[j].elem_begin + sj[j].elems) ++j;
        const a2x_unpack_job& J = sj[j];
        const long long e = i - J.elem_begin;
        if (J.kind == 2) {  // double sums -> float vector (bias gradients)
            J.dst[e] = (float)reinterpret_cast<const double*>(J.src)[J.row0 + e];
        } else if (J.kind == 0) {  // [tap][cout_pad][cin] rows [row0, row0 + rows) -> OIHW
            const int cin = J.b, kk = J.kk;
            const int tap = (int)(e % kk);
            const int ci = (int)((e / kk) % cin);
            const int co = (int)(e / ((long long)kk * cin));
            J.dst[e] = reinterpret_cast<const float*>(J.src)[((long long)tap * J.cout_pad + J.row0 + co) * cin + ci];
        } else {  // [ij][ci][co] -> [ci][co][i][j]
            const int cin = J.a, cout = J.b, ss = J.kk;
            const int ij = (int)(e % ss);
            const int co = (int)((e / ss) % cout);
            const int ci = (int)(e / ((long long)ss * cout));
            J.dst[e] = reinterpret_cast<const float*>(J.src)[((long long)ij * cin + ci) * cout + co];
        }
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

static int conv2d_fwd_impl(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                           const float* scale, const float* shift, int relu, int accumulate, double* stats,
                           const float* residual, unsigned long long seed, unsigned int site, float drop_p,
                           long long elem0, a2x_stream_t stream);

extern "C" {

int a2x_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int cout_pad, float* w_fwd, void* w_fwd16,
                         float* w_dgrad, void* w_dgrad16, a2x_stream_t stream) {
    A2X_REQUIRE(w_oihw && cout > 0 && cin > 0 && (ksize == 1 || ksize == 3 || ksize == 7) && cout_pad >= cout, "bad pack args");
    const int kk = ksize * ksize;
    pack_conv_w_kernel<<<grid_for((long long)kk * cout_pad * cin), 256, 0, (cudaStream_t)stream>>>(
        w_oihw, cout, cin, kk, cout_pad, w_fwd, (__nv_bfloat16*)w_fwd16, w_dgrad, (__nv_bfloat16*)w_dgrad16);
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

int a2x_pack_deconv_weight(const float* w_iohw, int cin, int cout, int s, float* w_fwd, void* w_fwd16, float* w_dgrad,
                           void* w_dgrad16, a2x_stream_t stream) {
    A2X_REQUIRE(w_iohw && cin > 0 && cout > 0 && s > 0, "bad pack args");
    pack_deconv_w_kernel<<<grid_for((long long)cin * cout * s * s), 256, 0, (cudaStream_t)stream>>>(
        w_iohw, cin, cout, s, w_fwd, (__nv_bfloat16*)w_fwd16, w_dgrad, (__nv_bfloat16*)w_dgrad16);
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

int a2x_pack_weights_batched(const a2x_pack_job* jobs_dev, int njobs, long long total_elems, a2x_stream_t stream) {
    A2X_REQUIRE(jobs_dev && njobs > 0 && njobs <= 128 && total_elems > 0, "pack_weights_batched: bad args (<= 128 jobs)");
    pack_batched_kernel<<<grid_for(total_elems), 256, 0, (cudaStream_t)stream>>>(jobs_dev, njobs);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_unpack_wgrads_batched(const a2x_unpack_job* jobs_dev, int njobs, long long total_elems, a2x_stream_t stream) {
    A2X_REQUIRE(jobs_dev && njobs > 0 && njobs <= 128 && total_elems > 0, "unpack_wgrads_batched: bad args (<= 128 jobs)");
    unpack_batched_kernel<<<grid_for(total_elems), 256, 0, (cudaStream_t)stream>>>(jobs_dev, njobs);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_conv2d_fwd(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                   const float* scale, const float* shift, int relu, double* stats, a2x_stream_t stream) {
    return a2x_conv2d_fwd_ex(s, x, w, y, scale, shift, relu, 0, stats, stream);
}

int a2x_conv2d_fwd_ex(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                      const float* scale, const float* shift, int relu, int accumulate, double* stats,
                      a2x_stream_t stream) {
    return conv2d_fwd_impl(s, x, w, y, scale, shift, relu, accumulate, stats, nullptr, 0ull, 0u, 0.f, 0, stream);
}

int a2x_linear_dropout_residual_fwd(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, float* y, int y_cs,
                                    const float* bias, const float* residual, unsigned long long seed, unsigned int site,
                                    float p, long long elem_offset, a2x_stream_t stream) {
    A2X_REQUIRE(s && s->ksize == 1 && s->stride == 1 && y && y_cs == s->cout && p >= 0.f && p < 1.f && elem_offset >= 0 &&
                    elem_offset % 8 == 0,
                "linear_dropout_residual_fwd: a token-wise linear (ksize 1, stride 1) with a dense fp32 output, 0 <= p < 1, "
                "elem_offset a multiple of 8");
    a2x_output out{};
    out.hi = y;
    out.cs = y_cs;
    return conv2d_fwd_impl(s, x, w, &out, nullptr, bias, 0, 0, nullptr, residual, seed, site, p, elem_offset, stream);
}

}  // extern "C"

static int conv2d_fwd_impl(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                           const float* scale, const float* shift, int relu, int accumulate, double* stats,
                           const float* residual, unsigned long long seed, unsigned int site, float drop_p,
                           long long elem0, a2x_stream_t stream) {
    if (int r = check_shape(s, false)) return r;
    A2X_REQUIRE(relu >= 0 && relu <= 2, "conv2d_fwd: activation must be 0 (none), 1 (ReLU) or 2 (GELU)");
    if (int r = check_operand(x, s->cin, "conv2d_fwd x")) return r;
    A2X_REQUIRE(w && w->w32 && y && (y->hi || y->b16) && (y->hi || !accumulate) && y->cs >= s->cout && y->cs % 4 == 0,
                "conv2d_fwd: bad weights/output");
    A2X_REQUIRE(!x->b16 || w->w16, "conv2d_fwd: split input needs bf16 weight planes");
    const int ho = (s->h - 1) / s->stride + 1, wo = (s->w - 1) / s->stride + 1;
    const int kk = s->ksize * s->ksize;
    A2X_REQUIRE(!stats || (!scale && !shift && !relu && !accumulate),
                "conv2d_fwd: fused statistics are of the raw conv output");
    if (s->ksize == 7) {
        // 7x7 stride-2 padding-3 stem (BevEncode.conv1, sub_modules/lss_submodule.py:318): input row 2i + r - 3 = 2 (i + a) + hp
        // is row i + a of parity view hp, a in {-2..1}: 49 taps over the four parity views of the stride-2 path, run as
        // groups of 9 taps (x 3 split products = the kernel's tap table) accumulating into the fp32 output
        A2X_REQUIRE(!scale && !shift && !relu && !accumulate && !stats && !residual && drop_p == 0.f && y->hi && !y->b16,
                    "conv2d_fwd: the 7x7 stem writes the raw fp32 convolution only");
        TgTap all[49];
        int nt = 0;
        for (int r = 0; r < 7; ++r)
            for (int cc = 0; cc < 7; ++cc) {
                const int oh = r - 3, ow = cc - 3;
                const int hp = oh & 1, wp = ow & 1;            // parity of the input row / column (two's complement: -3 & 1 = 1)
                TgTap t{};
                t.map = (int16_t)(hp * 2 + wp);
                t.dh = (int16_t)((oh - hp) / 2);
                t.dw = (int16_t)((ow - wp) / 2);
                t.dx = 0;
                t.btap = r * 7 + cc;
                all[nt++] = t;
            }
        for (int g0 = 0; g0 < 49; g0 += 9) {
            TgParams p{};
            p.tw_log2 = pick_tw_log2(ho, wo, TG_BM, 4, 7);
            const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
            if (int r = build_fwd_maps(s, x, p.amap, TW, TH, 0, p.amap + 4, p.amap + 8)) return r;
            const int ng = 49 - g0 < 9 ? 49 - g0 : 9;
            for (int i = 0; i < ng; ++i) p.taps[i] = all[g0 + i];
            p.ntaps = ng;
            const int bn = bn_for(s->cout);
            if (int r = make_w_map(&p.bmap, w->w32, kk, s->cout, s->cin, bn)) return r;
            if (x->b16) {
                p.ntaps = expand_split_taps(p.taps, p.ntaps, 4, kk);
                if (int r = make_w_map(&p.bmap16, w->w16, 2 * kk, s->cout, s->cin, bn, 1)) return r;
            }
            p.kchunks32 = s->cin / 32;
            p.kchunks16 = s->cin / 64;
            set_plain_out(p, y, ho, wo, 1, 0, 0);
            p.accumulate = g0 > 0;
            p.stat_c = s->cout;
            if (int r = run_tg(p, ho, wo, s->n, s->cout, (cudaStream_t)stream)) return r;
        }
        return 0;
    }
    A2X_REQUIRE((!residual && drop_p == 0.f) || (s->ksize == 1 && !y->b16 && !relu && !accumulate),
                "conv2d_fwd: the dropout / residual epilogue is the fp32-only token-wise linear one");
    if (s->ksize == 3 && s->stride == 1 && g_debug[7] != 1)
        return run_halo(x, s->n, s->h, s->w, s->cin, s->cout, w, +1, y, scale, shift, relu, accumulate, stats,
                        (cudaStream_t)stream);
    TgParams p{};
    p.tw_log2 = pick_tw_log2(ho, wo, TG_BM, 4, 7);
    const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
    const int nmaps = s->stride == 1 ? 1 : 4;
    if (int r = build_fwd_maps(s, x, p.amap, TW, TH, 0, p.amap + nmaps, p.amap + 2 * nmaps)) return r;
    p.ntaps = build_fwd_taps(s, p.taps);
    const int bn = bn_for(s->cout);
    if (int r = make_w_map(&p.bmap, w->w32, kk, s->cout, s->cin, bn)) return r;
    if (x->b16) {
        p.ntaps = expand_split_taps(p.taps, p.ntaps, nmaps, kk);
        if (int r = make_w_map(&p.bmap16, w->w16, 2 * kk, s->cout, s->cin, bn, 1)) return r;
    }
    p.kchunks32 = s->cin / 32;
    p.kchunks16 = s->cin / 64;
    set_plain_out(p, y, ho, wo, 1, 0, 0);
    p.scale = scale;
    p.shift = shift;
    p.relu = relu;
    if (g_debug[9]) p.relu = 77;  // debug: skip the epilogue stores (epilogue cost experiments)
    p.accumulate = accumulate;
    p.stats = stats;
    p.stat_c = s->cout;
    p.res = residual;
    p.drop.seed = seed;
    p.drop_elem0 = elem0;
    p.drop.site = site;
    p.drop.thresh = drop_p > 0.f ? (uint32_t)(drop_p * 65536.0f + 0.5f) : 0u;   // same rounding as dropout.cu
    p.drop.scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    return run_tg(p, ho, wo, s->n, s->cout, (cudaStream_t)stream);
}

extern "C" {

int a2x_conv2d_dgrad(const a2x_conv_shape* s, const a2x_operand* dy, const a2x_weights* w, float* dx, int dx_cs,
                     int accumulate, a2x_stream_t stream) {
    return a2x_conv2d_dgrad_ex(s, dy, w, dx, dx_cs, accumulate, nullptr, stream);
}

int a2x_conv2d_dgrad_ex(const a2x_conv_shape* s, const a2x_operand* dy, const a2x_weights* w, float* dx, int dx_cs,
                        int accumulate, const a2x_bn_bwd_stats* bn_stats, a2x_stream_t stream) {
    if (int r = check_shape(s, false)) return r;
    A2X_REQUIRE(!bn_stats || (s->ksize == 3 && s->stride == 1),
                "conv2d_dgrad: the fused BN backward reduction is implemented for 3x3 stride-1 windows");
    if (int r = check_operand(dy, s->cout, "conv2d_dgrad dy")) return r;
    A2X_REQUIRE(w && w->w32 && dx && dx_cs >= s->cin && dx_cs % 4 == 0, "conv2d_dgrad: bad weights/output");
    A2X_REQUIRE(!dy->b16 || w->w16, "conv2d_dgrad: split input needs bf16 weight planes");
    const int ho = (s->h - 1) / s->stride + 1, wo = (s->w - 1) / s->stride + 1;
    const int kk = s->ksize * s->ksize;
    const int n_class = s->stride == 1 ? 1 : 4;
    a2x_output out{dx, nullptr, 0, dx_cs};
    if (s->ksize == 3 && s->stride == 1 && g_debug[7] != 1)
        return run_halo(dy, s->n, s->h, s->w, s->cout, s->cin, w, -1, &out, nullptr, nullptr, 0, accumulate, nullptr,
                        (cudaStream_t)stream, bn_stats);
    A2X_REQUIRE(!bn_stats, "conv2d_dgrad: fused BN backward reduction needs the halo path");
    for (int cls = 0; cls < n_class; ++cls) {
        const int hp = cls >> 1, wp = cls & 1;
        const int step = s->stride;
        const int gh = (s->h - hp + step - 1) / step, gw = (s->w - wp + step - 1) / step;
        if (gh <= 0 || gw <= 0) continue;
        TgParams p{};
        p.tw_log2 = pick_tw_log2(gh, gw, TG_BM, 4, 7);
        const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
        if (int r = make_act_map(&p.amap[0], dy->hi, s->n, ho, wo, s->cout, dy->cs, 1, 0, 0, TW, TH)) return r;
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
        const int bn = bn_for(s->cin);
        if (int r = make_w_map(&p.bmap, w->w32, kk, s->cin, s->cout, bn)) return r;
        if (dy->b16) {
            const __nv_bfloat16* b = (const __nv_bfloat16*)dy->b16;
            if (int r = make_act_map(&p.amap[1], b, s->n, ho, wo, s->cout, dy->cs, 1, 0, 0, TW, TH, 0, 1)) return r;
            if (int r = make_act_map(&p.amap[2], b + dy->b16_plane, s->n, ho, wo, s->cout, dy->cs, 1, 0, 0, TW, TH, 0, 1))
                return r;
            p.ntaps = expand_split_taps(p.taps, p.ntaps, 1, kk);
            if (int r = make_w_map(&p.bmap16, w->w16, 2 * kk, s->cin, s->cout, bn, 1)) return r;
        }
        p.kchunks32 = s->cout / 32;
        p.kchunks16 = s->cout / 64;
        set_plain_out(p, &out, s->h, s->w, step, hp, wp);
        p.accumulate = accumulate;
        if (int r = run_tg(p, gh, gw, s->n, s->cin, (cudaStream_t)stream)) return r;
    }
    return 0;
}

// shared by conv and deconv weight gradients: fill the A-side maps of `a` (unshifted operand)
static int wg_a_maps(WgParams& p, const a2x_operand* a, int n, int h, int w, int c, int TW, int TH) {
    if (int r = make_act_map(&p.amap, a->hi, n, h, w, c, a->cs, 1, 0, 0, TW, TH, 1)) return r;
    if (a->b16) {
        const __nv_bfloat16* b = (const __nv_bfloat16*)a->b16;
        if (int r = make_act_map(&p.amap16[0], b, n, h, w, c, a->cs, 1, 0, 0, TW, TH, 0, 1)) return r;
        if (int r = make_act_map(&p.amap16[1], b + a->b16_plane, n, h, w, c, a->cs, 1, 0, 0, TW, TH, 0, 1)) return r;
    }
    return 0;
}

int a2x_conv2d_wgrad(const a2x_conv_shape* s, const a2x_operand* x, const a2x_operand* dy, float* dw_packed,
                     a2x_stream_t stream) {
    if (int r = check_shape(s, false)) return r;
    if (int r = check_operand(x, s->cin, "conv2d_wgrad x")) return r;
    if (int r = check_operand(dy, s->cout, "conv2d_wgrad dy")) return r;
    A2X_REQUIRE(dw_packed, "conv2d_wgrad: null output");
    const bool split = x->b16 && dy->b16;
    const int ho = (s->h - 1) / s->stride + 1, wo = (s->w - 1) / s->stride + 1;
    if (split && s->ksize == 3 && s->stride == 1 && g_debug[7] != 1)
        return run_wr(s, x, dy, dw_packed, (cudaStream_t)stream);
    WgParams p{};
    p.tw_log2 = pick_tw_log2(ho, wo, WG_PIX, 2, 5);
    const int TW = 1 << p.tw_log2, TH = WG_PIX >> p.tw_log2;
    a2x_operand dyo = *dy, xo = *x;
    if (!split) dyo.b16 = xo.b16 = nullptr;
    if (int r = wg_a_maps(p, &dyo, s->n, ho, wo, s->cout, TW, TH)) return r;
    if (int r = build_fwd_maps(s, &xo, p.bmap, TW, TH, 1, p.bmap16[0], p.bmap16[1])) return r;
    p.ntaps = build_fwd_taps(s, p.taps);
    p.ca = s->cout;
    p.cb = s->cin;
    p.dw = dw_packed;
    return run_wg(p, ho, wo, s->n, split, (cudaStream_t)stream);
}

int a2x_deconv_fwd(const a2x_conv_shape* s, const a2x_operand* x, const a2x_weights* w, const a2x_output* y,
                   const float* scale, const float* shift, int relu, double* stats, a2x_stream_t stream) {
    if (int r = check_shape(s, true)) return r;
    if (int r = check_operand(x, s->cin, "deconv_fwd x")) return r;
    A2X_REQUIRE(w && w->w32 && y && (y->hi || y->b16) && y->cs >= s->cout && y->cs % 4 == 0, "deconv_fwd: bad weights/output");
    A2X_REQUIRE(!x->b16 || w->w16, "deconv_fwd: split input needs bf16 weight planes");
    const int st = s->stride;
    TgParams p{};
    p.tw_log2 = pick_tw_log2(s->h, s->w, TG_BM, 4, 7);
    const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
    if (int r = make_act_map(&p.amap[0], x->hi, s->n, s->h, s->w, s->cin, x->cs, 1, 0, 0, TW, TH)) return r;
    p.ntaps = 1;
    p.taps[0] = TgTap{0, 0, 0, 0, 0, 0};
    const int ncols = st * st * s->cout;
    const int bn = bn_for(ncols);
    if (int r = make_w_map(&p.bmap, w->w32, 1, ncols, s->cin, bn)) return r;
    if (x->b16) {
        const __nv_bfloat16* b = (const __nv_bfloat16*)x->b16;
        if (int r = make_act_map(&p.amap[1], b, s->n, s->h, s->w, s->cin, x->cs, 1, 0, 0, TW, TH, 0, 1)) return r;
        if (int r = make_act_map(&p.amap[2], b + x->b16_plane, s->n, s->h, s->w, s->cin, x->cs, 1, 0, 0, TW, TH, 0, 1))
            return r;
        p.ntaps = expand_split_taps(p.taps, 1, 1, 1);
        if (int r = make_w_map(&p.bmap16, w->w16, 2, ncols, s->cin, bn, 1)) return r;
    }
    p.kchunks32 = s->cin / 32;
    p.kchunks16 = s->cin / 64;
    const long long W2 = (long long)s->w * st;
    p.out.hi = y->hi;
    p.out.b16 = (__nv_bfloat16*)y->b16;
    p.out.ps = y->b16_plane;
    p.osn = (long long)s->h * st * W2 * y->cs;
    p.osh = (long long)st * W2 * y->cs;
    p.osw = (long long)st * y->cs;
    p.sub_c = s->cout;
    p.sub_s = st;
    p.sub_sh = W2 * y->cs;
    p.sub_sw = y->cs;
    p.scale = scale;
    p.shift = shift;
    p.relu = relu;
    p.accumulate = 0;
    p.stats = stats;
    p.stat_c = s->cout;
    A2X_REQUIRE(!stats || (!scale && !shift && !relu), "deconv_fwd: fused statistics are of the raw output");
    return run_tg(p, s->h, s->w, s->n, ncols, (cudaStream_t)stream);
}

int a2x_deconv_dgrad(const a2x_conv_shape* s, const a2x_operand* dy, const a2x_weights* w, float* dx, int dx_cs,
                     int accumulate, a2x_stream_t stream) {
    if (int r = check_shape(s, true)) return r;
    if (int r = check_operand(dy, s->cout, "deconv_dgrad dy")) return r;
    A2X_REQUIRE(w && w->w32 && dx && dx_cs >= s->cin && dx_cs % 4 == 0, "deconv_dgrad: bad weights/output");
    A2X_REQUIRE(!dy->b16 || w->w16, "deconv_dgrad: split input needs bf16 weight planes");
    const int st = s->stride;
    a2x_output out{dx, nullptr, 0, dx_cs};
    // one launch per sub-row i; its s sub-columns j are the taps, each through its own strided view of dy
    for (int i = 0; i < st; ++i) {
        TgParams p{};
        p.tw_log2 = pick_tw_log2(s->h, s->w, TG_BM, 4, 7);
        const int TW = 1 << p.tw_log2, TH = TG_BM >> p.tw_log2;
        p.ntaps = 0;
        for (int j = 0; j < st; ++j) {
            if (int r = make_act_map(&p.amap[j], dy->hi, s->n, s->h * st, s->w * st, s->cout, dy->cs, st, i, j, TW, TH))
                return r;
            if (dy->b16) {
                const __nv_bfloat16* b = (const __nv_bfloat16*)dy->b16;
                if (int r = make_act_map(&p.amap[st + j], b, s->n, s->h * st, s->w * st, s->cout, dy->cs, st, i, j, TW,
                                         TH, 0, 1))
                    return r;
                if (int r = make_act_map(&p.amap[2 * st + j], b + dy->b16_plane, s->n, s->h * st, s->w * st, s->cout,
                                         dy->cs, st, i, j, TW, TH, 0, 1))
                    return r;
            }
            TgTap t{};
            t.map = (int16_t)j;
            t.btap = i * st + j;
            p.taps[p.ntaps++] = t;
        }
        const int bn = bn_for(s->cin);
        if (int r = make_w_map(&p.bmap, w->w32, st * st, s->cin, s->cout, bn)) return r;
        if (dy->b16) {
            p.ntaps = expand_split_taps(p.taps, p.ntaps, st, st * st);
            if (int r = make_w_map(&p.bmap16, w->w16, 2 * st * st, s->cin, s->cout, bn, 1)) return r;
        }
        p.kchunks32 = s->cout / 32;
        p.kchunks16 = s->cout / 64;
        set_plain_out(p, &out, s->h, s->w, 1, 0, 0);
        p.accumulate = (i > 0) ? 1 : accumulate;
        if (int r = run_tg(p, s->h, s->w, s->n, s->cin, (cudaStream_t)stream)) return r;
    }
    return 0;
}

int a2x_deconv_wgrad(const a2x_conv_shape* s, const a2x_operand* x, const a2x_operand* dy, float* dw_packed,
                     a2x_stream_t stream) {
    if (int r = check_shape(s, true)) return r;
    if (int r = check_operand(x, s->cin, "deconv_wgrad x")) return r;
    if (int r = check_operand(dy, s->cout, "deconv_wgrad dy")) return r;
    A2X_REQUIRE(dw_packed, "deconv_wgrad: null output");
    const bool split = x->b16 && dy->b16;
    const int st = s->stride;
    a2x_operand xo = *x;
    if (!split) xo.b16 = nullptr;
    for (int i = 0; i < st; ++i) {
        WgParams p{};
        p.tw_log2 = pick_tw_log2(s->h, s->w, WG_PIX, 2, 5);
        const int TW = 1 << p.tw_log2, TH = WG_PIX >> p.tw_log2;
        if (int r = wg_a_maps(p, &xo, s->n, s->h, s->w, s->cin, TW, TH)) return r;
        p.ntaps = 0;
        for (int j = 0; j < st; ++j) {
            if (int r = make_act_map(&p.bmap[j], dy->hi, s->n, s->h * st, s->w * st, s->cout, dy->cs, st, i, j, TW, TH, 1))
                return r;
            if (split) {
                const __nv_bfloat16* b = (const __nv_bfloat16*)dy->b16;
                if (int r = make_act_map(&p.bmap16[0][j], b, s->n, s->h * st, s->w * st, s->cout, dy->cs, st, i, j, TW,
                                         TH, 0, 1))
                    return r;
                if (int r = make_act_map(&p.bmap16[1][j], b + dy->b16_plane, s->n, s->h * st, s->w * st, s->cout, dy->cs,
                                         st, i, j, TW, TH, 0, 1))
                    return r;
            }
            TgTap t{};
            t.map = (int16_t)j;
            p.taps[p.ntaps++] = t;
        }
        p.ca = s->cin;
        p.cb = s->cout;
        p.dw = dw_packed + (long long)i * st * s->cin * s->cout;  // [(i, j)][ci][co]
        if (int r = run_wg(p, s->h, s->w, s->n, split, (cudaStream_t)stream)) return r;
    }
    return 0;
}

}  // extern "C"
