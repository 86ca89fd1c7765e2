// Halo variant of the tap-GEMM for 3x3 stride-1 windows (conv forward and data gradient): the 9 window taps read
// overlapping pixels, so instead of 9 TMA loads of a 128-pixel A tile per k-chunk, ONE haloed tile
// (18 x 10 pixels around a 16 x 8 output tile) is loaded and the nine tcgen05.mma groups address shifted windows of
// it through the smem descriptor alone:  start = slot + ((1+dh)*10 + (1+dw)) * 128 B, stride between 8-row groups
// (SBO) = 10 * 128 B (one halo row). A traffic from L2 drops 6.25x (23 KB instead of 144 KB per k-chunk); the B
// (weight) tiles stream through their own ring. Everything else (TMEM accumulator, split-precision passes, epilogue
// with fused BN statistics) is shared with tapgemm.cuh.
#pragma once
#include "tapgemm.cuh"

namespace a2x {

constexpr int TH_ROWS = 18, TH_COLS = 10;               // haloed tile: (16 + 2) x (8 + 2) pixels
constexpr int TH_A_BYTES = TH_ROWS * TH_COLS * 128;     // 23040
constexpr int TH_A_SLOT = 23 * 1024;                    // 1024-aligned slot
constexpr int TH_NA = 2;                                // A slots

struct ThParams {
    CUtensorMap amap[2];  // halo-box views in pass order: split = {h16, l16} (bf16); single plane = {fp32}
    CUtensorMap bmap;     // weights: split = bf16 [18][cols][k] (9 h16 taps, then 9 l16 taps); single = fp32 [9][cols][k]
    int npass;            // 1 (single plane) or 2 (split: A_h16 x 18 weight taps, A_l16 x the 9 h16 taps)
    int pass_taps[2];     // weight taps per pass (tap t uses window offset t % 9)
    int kind;             // 0 = tf32 (32 channels per 128-byte row), 1 = bf16 (64 channels)
    int kchunks;          // K / (32 | 64)
    int8_t dh[9], dw[9];  // window offsets per tap (forward: r-1, c-1; data gradient: 1-r, 1-c)
    int n_img, gh, gw, tiles_h, tiles_w;
    SplitOut out;
    long long osn, osh, osw;
    int ncols;
    const float* scale;
    const float* shift;
    int relu, accumulate;
    double* stats;
    int stat_c;
    const float* bz;  // fused BN+ReLU backward reduction (see TgEpi); null = plain statistics
    const float* bscale;
    const float* bshift;
    const float* bmean;
    const float* binvstd;
    int base_offset_mode;  // debug: 0 = descriptor base_offset 0, 1 = (start >> 7) & 7
};

template <int BN, int NB>
struct ThSmem {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int OFF_B = TH_NA * TH_A_SLOT;
    static constexpr int BAR_OFF = OFF_B + NB * B_BYTES;
    static constexpr int NBAR = 2 * TH_NA + 2 * NB + 4;
    static constexpr int STAT_OFF = BAR_OFF + NBAR * 8 + 16;
    static constexpr int XP_OFF = STAT_OFF + 4 * 2 * BN * 4;
    static constexpr int TOTAL = XP_OFF + 4 * TG_XP_BYTES + 1024;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                     : (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int NB>
__global__ void __launch_bounds__(192) tapgemm_halo_kernel(const __grid_constant__ ThParams p) {
    using L = ThSmem<BN, NB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* a_empty = a_full + TH_NA;
    uint64_t* b_full = a_empty + TH_NA;
    uint64_t* b_empty = b_full + NB;
    uint64_t* tmem_full = b_empty + NB;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* sstat = reinterpret_cast<float*>(smem + L::STAT_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_m = p.n_img * p.tiles_h * p.tiles_w;
    const int tiles_n = (p.ncols + BN - 1) / BN;
    const int n_tiles = tiles_m * tiles_n;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TH_NA; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<L::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int t = tile % tiles_m;
                const int n0 = (tile / tiles_m) * BN;
                const int tw_i = t % p.tiles_w;
                t /= p.tiles_w;
                const int th_i = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th_i * 16, w0 = tw_i * 8;
                const int kw = p.kind ? 64 : 32;
                for (int pass = 0; pass < p.npass; ++pass) {
                    const int nt = p.pass_taps[pass];
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        mbar_arrive_expect_tx(&a_full[sa], TH_A_BYTES);
                        tma_load_5d(smem + sa * TH_A_SLOT, &p.amap[pass], &a_full[sa], kc * kw, w0 - 1, 0, h0 - 1, img);
                        if (++sa == TH_NA) {
                            sa = 0;
                            pa ^= 1;
                        }
                        for (int tap = 0; tap < nt; ++tap) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            mbar_arrive_expect_tx(&b_full[sb], L::B_BYTES);
                            tma_load_3d(smem + L::OFF_B + sb * L::B_BYTES, &p.bmap, &b_full[sb], kc * kw, n0, tap);
                            if (++sb == NB) {
                                sb = 0;
                                pb ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc32 = make_idesc_tf32(TG_BM, BN, 0, 0);
            constexpr uint32_t idesc16 = make_idesc_bf16(TG_BM, BN, 0, 0);
            constexpr uint32_t ahi = desc_hi_word(TH_COLS * 128, 2);  // 8-row groups one halo row (1280 B) apart
            constexpr uint32_t bhi = desc_hi_word(1024, 2);
            uint32_t a_tap[9];  // window offsets (>> 4) of the nine taps inside a halo slot
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
                a_tap[tap] = (uint32_t)(((1 + p.dh[tap]) * TH_COLS + (1 + p.dw[tap])) * 128) >> 4;
            const uint32_t a_lo0 = desc_lo_word(smem_u32(smem), 16);
            const uint32_t b_lo0 = desc_lo_word(smem_u32(smem) + L::OFF_B, 16);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            uint32_t sblo = 0;  // (sb * B_BYTES) >> 4
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t tacc = tmem_base + buf * BN;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                uint32_t acc = 0;
                const int kind = p.kind;
                for (int pass = 0; pass < p.npass; ++pass) {
                    const int ngroups = p.pass_taps[pass] / 9;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&a_full[sa], pa);
                        tc_fence_after();
                        const uint32_t aslot = a_lo0 + sa * (TH_A_SLOT >> 4);
                        for (int g = 0; g < ngroups; ++g) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                mbar_wait(&b_full[sb], pb);
                                tc_fence_after();
                                const uint32_t alo = aslot + a_tap[tap], blo = b_lo0 + sblo;
                                if (kind) {
                                    umma_bf16_lh(tacc, alo, ahi, blo, bhi, idesc16, acc);
                                    umma_bf16_lh(tacc, alo + 2, ahi, blo + 2, bhi, idesc16, 1);
                                    umma_bf16_lh(tacc, alo + 4, ahi, blo + 4, bhi, idesc16, 1);
                                    umma_bf16_lh(tacc, alo + 6, ahi, blo + 6, bhi, idesc16, 1);
                                } else {
                                    umma_tf32_lh(tacc, alo, ahi, blo, bhi, idesc32, acc);
                                    umma_tf32_lh(tacc, alo + 2, ahi, blo + 2, bhi, idesc32, 1);
                                    umma_tf32_lh(tacc, alo + 4, ahi, blo + 4, bhi, idesc32, 1);
                                    umma_tf32_lh(tacc, alo + 6, ahi, blo + 6, bhi, idesc32, 1);
                                }
                                acc = 1;
                                umma_commit(&b_empty[sb]);
                                sblo += L::B_BYTES >> 4;
                                if (++sb == NB) {
                                    sb = 0;
                                    pb ^= 1;
                                    sblo = 0;
                                }
                            }
                        }
                        umma_commit(&a_empty[sa]);  // the halo tile is free once its taps retire
                        if (++sa == TH_NA) {
                            sa = 0;
                            pa ^= 1;
                        }
                    }
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int et = threadIdx.x - 64;
        TgEpi e;
        e.out = p.out; e.osn = p.osn; e.osh = p.osh; e.osw = p.osw; e.sub_c = 1 << 30; e.sub_s = 1;
        e.sub_sh = 0; e.sub_sw = 0; e.ncols = p.ncols; e.scale = p.scale; e.shift = p.shift;
        e.relu = p.relu; e.accumulate = p.accumulate; e.stats = p.stats; e.stat_c = p.stat_c; e.gh = p.gh; e.gw = p.gw;
        e.res = nullptr; e.drop.thresh = 0; e.drop_elem0 = 0;
        e.bz = p.bz; e.bscale = p.bscale; e.bshift = p.bshift; e.bmean = p.bmean; e.binvstd = p.binvstd;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            int t = tile % tiles_m;
            const int n0 = (tile / tiles_m) * BN;
            const int tw_i = t % p.tiles_w;
            t /= p.tiles_w;
            const int th_i = t % p.tiles_h;
            const int img = t / p.tiles_h;
            const int buf = it & 1;
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            tg_epilogue<BN>(e, tmem_base + buf * BN, q, lane, et, img, th_i * 16, tw_i * 8, 3, n0, sstat,
                            reinterpret_cast<float*>(smem + L::XP_OFF + q * TG_XP_BYTES));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<L::TMEM_COLS>(tmem_base);
}

}  // namespace a2x
