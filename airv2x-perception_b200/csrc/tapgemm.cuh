// Tap-GEMM: the one tensor-core kernel family behind every dense contraction on the BEV path.
//
//   D[pixel, col] = sum_{tap} sum_{k} A_tap[pixel + shift(tap), k] * B[tap][col, k]
//
// A is an NHWC activation tensor seen through rank-5 TMA tensor maps (c, w, x, h, n); TMA's out-of-bounds zero
// fill implements the conv zero padding, per-tap coordinate shifts implement the 3x3 window, and "parity view" maps
// (same memory, doubled strides) implement stride-2 convs and their data gradients. B is a packed weight tensor
// [tap][col][k] (K-major). One CTA computes a 128-pixel x BN-column tile: a TMA producer lane fills a STAGES-deep
// smem ring (128B-swizzled), one elected lane issues tcgen05.mma into a TMEM accumulator, four epilogue warps drain
// it with tcgen05.ld, apply the per-channel affine / ReLU, optionally reduce the BatchNorm batch statistics, and
// store NHWC rows (each thread owns one pixel => full 128-byte lines).
//
// Precision ("split" mode): every operand v is carried as hi = tf32_rn(v) (fp32) plus a bf16 pair
// (h16, l16) = (bf16(hi), bf16(v - hi)). Each window tap expands to three tap entries accumulating into the SAME fp32
// TMEM tile: hi*hi with kind::tf32 (K = 8 / instruction), l16*h16 and h16*l16 with kind::f16 bf16 (K = 16 /
// instruction, i.e. twice the rate and half the bytes). Error budget ~2^-20 relative per product.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.ConvTranspose2d on the reference path
// (opencood/models/common_modules/base_bev_backbone.py:41-105, downsample_conv.py:18-32,
//  airv2x_where2com.py:59-69).
#pragma once
#include "a2x_ptx.cuh"

namespace a2x {

constexpr int TG_BM = 128;  // pixels per tile (UMMA M)
constexpr int TG_A_BYTES = TG_BM * 128;
constexpr int TG_MAX_TAPS = 27;  // 9 window taps x 3 split products
constexpr int TG_MAX_MAPS = 12;  // 4 parity views x {hi fp32, h16, l16}

struct TgTap {
    int16_t map;   // which A tensor map
    int16_t dw;    // shift along tensor-map dim 1 (w)
    int16_t dx;    // coordinate along dim 2 (x)
    int16_t dh;    // shift along dim 3 (h)
    int32_t btap;  // coordinate along B dim 2
    int32_t kind;  // 0 = tf32 (fp32 maps, 32 channels / 128-byte row), 1 = bf16 (64 channels / row)
};

struct TgParams {
    CUtensorMap amap[TG_MAX_MAPS];
    CUtensorMap bmap;    // fp32 weights  [taps][cols][k]       (hi plane)
    CUtensorMap bmap16;  // bf16 weights  [2*taps][cols][k]     (h16 taps, then l16 taps)
    TgTap taps[TG_MAX_TAPS];
    int ntaps;
    int kchunks32;  // K / 32 (tf32 taps)
    int kchunks16;  // K / 64 (bf16 taps)
    // GEMM pixel grid
    int n_img, gh, gw;
    int tw_log2;  // tile is (128 >> tw_log2) rows x (1 << tw_log2) cols
    int tiles_h, tiles_w;
    // output addressing (elements)
    SplitOut out;
    long long osn, osh, osw;
    int sub_c, sub_s;  // column -> (sub, c) split for deconv scatter; sub_c >= total cols otherwise
    long long sub_sh, sub_sw;
    int ncols;  // valid columns (multiple of 32)
    // epilogue
    const float* scale;  // per output channel (index = col % sub_c), may be null
    const float* shift;  // may be null
    int relu;
    int accumulate;  // out += result (single-plane outputs only)
    // fused BatchNorm batch statistics of the raw result: stats[ch] += sum, stats[stat_c + ch] += sum of squares
    double* stats;
    int stat_c;
};

template <int BN, int STAGES>
struct TgSmem {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = TG_A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int STAT_OFF = BAR_OFF + (2 * STAGES + 1) * 8 + 16;
    static constexpr int TOTAL = STAT_OFF + 4 * 2 * BN * 4 + 1024;  // [4 warps][2][BN] stats + alignment slack
};

// lane L ends up with the sum over the warp's 32 lanes of v[L] (31 shuffles); v is destroyed
__device__ __forceinline__ float warp_transpose_sum(float* v, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) tapgemm_kernel(const __grid_constant__ TgParams p) {
    using L = TgSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* sstat = reinterpret_cast<float*>(smem + L::STAT_OFF);  // [4 warps][2][BN] per-tile channel sums

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates
    int t = blockIdx.x;
    const int tw_i = t % p.tiles_w;
    t /= p.tiles_w;
    const int th_i = t % p.tiles_h;
    const int img = t / p.tiles_h;
    const int TW = 1 << p.tw_log2;
    const int TH = TG_BM >> p.tw_log2;
    const int h0 = th_i * TH, w0 = tw_i * TW;
    const int n0 = blockIdx.y * BN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&p.bmap);
            tma_prefetch_desc(&p.amap[0]);
            int stage = 0;
            uint32_t phase = 0;
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const TgTap tp = p.taps[tap];
                const CUtensorMap* am = &p.amap[tp.map];
                const CUtensorMap* bm = tp.kind ? &p.bmap16 : &p.bmap;
                const int nk = tp.kind ? p.kchunks16 : p.kchunks32;
                const int kw = tp.kind ? 64 : 32;  // channels per 128-byte row
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + TG_A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    tma_load_5d(sa, am, &full_bar[stage], kc * kw, w0 + tp.dw, tp.dx, h0 + tp.dh, img);
                    tma_load_3d(sb, bm, &full_bar[stage], kc * kw, n0, tp.btap);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc32 = make_idesc_tf32(TG_BM, BN, 0, 0);
            constexpr uint32_t idesc16 = make_idesc_bf16(TG_BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t first = 1;
            for (int tap = 0; tap < p.ntaps; ++tap) {
                const int kind = p.taps[tap].kind;
                const int nk = kind ? p.kchunks16 : p.kchunks32;
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t sb = sa + TG_A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per 128-byte row (8 tf32 or 16 bf16)
                        const uint64_t ad = make_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t bd = make_smem_desc_sw128(sb + k * 32, 16, 1024);
                        const uint32_t acc = (first && k == 0) ? 0u : 1u;
                        if (kind) umma_bf16(tmem_base, ad, bd, idesc16, acc);
                        else umma_tf32(tmem_base, ad, bd, idesc32, acc);
                    }
                    first = 0;
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            umma_commit(accum_bar);
        }
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrant (warp % 4)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int h = h0 + (row >> p.tw_log2);
        const int w = w0 + (row & (TW - 1));
        const bool valid = (h < p.gh) && (w < p.gw);
        const long long obase = (long long)img * p.osn + (long long)h * p.osh + (long long)w * p.osw;
        const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            const int c0 = n0 + j * 32;
            if (c0 >= p.ncols) break;
            float v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(j * 32), v);
            tmem_ld_wait();
            const int sub = c0 / p.sub_c;
            const int cc = c0 - sub * p.sub_c;
            if (p.scale != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= __ldg(p.scale + cc + i);
            }
            if (p.shift != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(p.shift + cc + i);
            }
            if (valid) {
                const long long off =
                    obase + (long long)(sub / p.sub_s) * p.sub_sh + (long long)(sub % p.sub_s) * p.sub_sw + cc;
                if (p.accumulate) {
                    const float4* o4 = reinterpret_cast<const float4*>(p.out.hi + off);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 prev = o4[i];
                        v[4 * i] += prev.x;
                        v[4 * i + 1] += prev.y;
                        v[4 * i + 2] += prev.z;
                        v[4 * i + 3] += prev.w;
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    store_split4(p.out, off + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
            if (p.stats != nullptr) {  // warp-uniform
                float sq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = valid ? v[i] : 0.f;
                    sq[i] = v[i] * v[i];
                }
                const float s1 = warp_transpose_sum(v, lane);
                const float s2 = warp_transpose_sum(sq, lane);
                sstat[(q * 2 + 0) * BN + j * 32 + lane] = s1;  // each (warp, column) written exactly once:
                sstat[(q * 2 + 1) * BN + j * 32 + lane] = s2;  // fixed-order combine below => deterministic
            }
        }
        if (p.stats != nullptr) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = et; i < BN; i += 128) {
                const int col = n0 + i;
                if (col < p.ncols) {
                    const int ch = col % p.sub_c;
                    const float t1 = (sstat[i] + sstat[2 * BN + i]) + (sstat[4 * BN + i] + sstat[6 * BN + i]);
                    const float t2 = (sstat[BN + i] + sstat[3 * BN + i]) + (sstat[5 * BN + i] + sstat[7 * BN + i]);
                    atomicAdd(&p.stats[ch], (double)t1);
                    atomicAdd(&p.stats[p.stat_c + ch], (double)t2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
    }
}

}  // namespace a2x
