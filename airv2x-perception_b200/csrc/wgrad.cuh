// Weight-gradient tap-GEMM (the K dimension is pixels):
//
//   dW[tap][a][b] = sum_{pixel} A[pixel, a] * B_tap[pixel + shift(tap), b]
//
// A (unshifted, e.g. dY) and B (shifted per tap through the same parity-view tensor maps the forward uses, e.g. X)
// are NHWC, so both UMMA operands are MN-major: a TMA box (one 128-byte channel row x 32 pixels) lands in smem as one
// swizzled "atom" [32 pixel rows][128 bytes]; M = 128 spans several atoms (LBO = atom size).
//   * single-plane mode (kind::tf32): atoms hold 32 fp32 channels; MN-major tf32 operands MUST use the 128B swizzle with
//     32-byte atomicity (UMMA layout SWIZZLE_128B_BASE32B, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 4-row groups,
//     SBO = 512, K = 8 pixels per MMA.
//   * split mode (bf16 x 3: h*h + l*h + h*l, kind::f16): atoms hold 64 bf16 channels, plain 128B swizzle, 8-row groups,
//     SBO = 1024, K = 16 pixels per MMA; only the two bf16 planes of each operand are loaded.
// All products accumulate into the same fp32 TMEM tile. A CTA owns one (128 x BN) x TPC-taps output block and a
// contiguous range of 32-pixel tiles (split-K); partial sums leave through vectorised red.global.add.f32.
//
// wgrad_row_kernel (below) is the fast path for 3x3 stride-1 windows in split mode: the pixel tile is a 32-pixel row
// segment and the three horizontally adjacent taps share ONE 34-pixel haloed load of X (their operands are the same
// smem rows shifted by one 128-byte row in the descriptor), which cuts the L2 -> smem traffic per MMA by ~2x against
// per-tap loads and lets a CTA keep three taps (3 x BN TMEM columns) in flight.
//
// Replaces autograd's cuDNN wgrad for nn.Conv2d / nn.ConvTranspose2d (reference: torch autograd over
// opencood/models/common_modules/base_bev_backbone.py:41-105, downsample_conv.py:18-32).
#pragma once
#include "a2x_ptx.cuh"
#include "tapgemm.cuh"

namespace a2x {

constexpr int WG_PIX = 32;                   // pixels per k-step
constexpr int WG_ATOM_BYTES = WG_PIX * 128;  // one (32 px x 128 B) atom
constexpr int WG_MAX_BMAPS = 4;              // parity views (conv stride 2) or sub-columns (deconv)

struct WgParams {
    CUtensorMap amap;                  // A hi (fp32)
    CUtensorMap amap16[2];             // A h16, l16 (bf16)
    CUtensorMap bmap[WG_MAX_BMAPS];    // B hi views
    CUtensorMap bmap16[2][WG_MAX_BMAPS];  // B h16 / l16 views
    TgTap taps[9];
    int ntaps;
    int ca, cb;  // channel counts (multiples of 32; of 64 in split mode)
    int n_img, tiles_h, tiles_w, tw_log2;  // 32-pixel tiles: (32 >> tw_log2) rows x (1 << tw_log2) cols
    int tiles_per_cta;
    int n_tiles_b;  // number of BN-wide column tiles
    float* dw;      // [ntaps][ca][cb], accumulated with red.add (caller zeroes)
    uint32_t lbo_bytes, sbo_bytes, layout;  // tf32 MN-major descriptor fields (debug-overridable)
};

template <int BN, int TPC, int STAGES, bool SPLIT>
struct WgSmem {
    static constexpr int A32 = SPLIT ? 0 : 4 * WG_ATOM_BYTES;                       // 128 ch fp32
    static constexpr int B32 = SPLIT ? 0 : TPC * (BN / 32) * WG_ATOM_BYTES;
    static constexpr int A16 = SPLIT ? 2 * WG_ATOM_BYTES : 0;                        // 128 ch bf16, per plane
    static constexpr int B16 = SPLIT ? TPC * (BN / 64) * WG_ATOM_BYTES : 0;          // per plane
    static constexpr int OFF_B32 = A32;
    static constexpr int OFF_A16 = OFF_B32 + B32;   // h16 then l16
    static constexpr int OFF_B16 = OFF_A16 + 2 * A16;
    static constexpr int STAGE_BYTES = OFF_B16 + 2 * B16;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
    static constexpr int TMEM_COLS = (TPC * BN <= 32) ? 32 : (TPC * BN <= 64) ? 64 : (TPC * BN <= 128) ? 128
                                     : (TPC * BN <= 256) ? 256 : 512;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

template <int BN, int TPC, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(192) wgrad_kernel(const __grid_constant__ WgParams p) {
    using L = WgSmem<BN, TPC, STAGES, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m0 = (blockIdx.y / p.n_tiles_b) * 128;
    const int n0 = (blockIdx.y % p.n_tiles_b) * BN;
    const int tap0 = blockIdx.z * TPC;
    const int a_atoms = min(4, (p.ca - m0) / 32);
    const int b_atoms = min(BN / 32, (p.cb - n0) / 32);
    const int total_tiles = p.n_img * p.tiles_h * p.tiles_w;
    const int tile_begin = blockIdx.x * p.tiles_per_cta;
    const int tile_end = min(total_tiles, tile_begin + p.tiles_per_cta);
    const int ntiles = tile_end - tile_begin;
    if (ntiles <= 0) return;  // uniform for the whole CTA

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<L::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const int TW = 1 << p.tw_log2;
            const int TH = WG_PIX >> p.tw_log2;
            const int a16 = a_atoms / 2, b16 = b_atoms / 2;  // 64-channel bf16 atoms
            const uint32_t tx_bytes =
                (SPLIT ? 2 * (a16 + TPC * b16) : (a_atoms + TPC * b_atoms)) * WG_ATOM_BYTES;
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                int t = tile;
                const int tw_i = t % p.tiles_w;
                t /= p.tiles_w;
                const int th_i = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th_i * TH, w0 = tw_i * TW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * L::STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                if (!SPLIT) {
                    for (int a = 0; a < a_atoms; ++a)
                        tma_load_5d(st + a * WG_ATOM_BYTES, &p.amap, &full_bar[stage], m0 + a * 32, w0, 0, h0, img);
                    for (int tt = 0; tt < TPC; ++tt) {
                        const TgTap tp = p.taps[tap0 + tt];
                        for (int b = 0; b < b_atoms; ++b)
                            tma_load_5d(st + L::OFF_B32 + (tt * (BN / 32) + b) * WG_ATOM_BYTES, &p.bmap[tp.map],
                                        &full_bar[stage], n0 + b * 32, w0 + tp.dw, tp.dx, h0 + tp.dh, img);
                    }
                }
                if (SPLIT) {
                    for (int pl = 0; pl < 2; ++pl) {
                        for (int a = 0; a < a16; ++a)
                            tma_load_5d(st + L::OFF_A16 + pl * L::A16 + a * WG_ATOM_BYTES, &p.amap16[pl],
                                        &full_bar[stage], m0 + a * 64, w0, 0, h0, img);
                        for (int tt = 0; tt < TPC; ++tt) {
                            const TgTap tp = p.taps[tap0 + tt];
                            for (int b = 0; b < b16; ++b)
                                tma_load_5d(st + L::OFF_B16 + pl * L::B16 + (tt * (BN / 64) + b) * WG_ATOM_BYTES,
                                            &p.bmap16[pl][tp.map], &full_bar[stage], n0 + b * 64, w0 + tp.dw, tp.dx,
                                            h0 + tp.dh, img);
                        }
                    }
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc32 = make_idesc_tf32(128, BN, 1, 1);
            constexpr uint32_t idesc16 = make_idesc_bf16(128, BN, 1, 1);
            // loop-invariant descriptor words (MN-major): tf32 atoms use the 32-byte-atomicity swizzle (layout 1,
            // SBO 512), bf16 atoms the plain 128B swizzle (layout 2, SBO 1024); LBO = atom size for both
            const uint32_t hi32 = desc_hi_word(p.sbo_bytes, p.layout);
            constexpr uint32_t hi16 = desc_hi_word(1024, 2);
            const uint32_t lo0 = desc_lo_word(smem_u32(smem), p.lbo_bytes);
            const uint32_t lo16 = desc_lo_word(smem_u32(smem), WG_ATOM_BYTES);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t slo = 0;  // (stage * STAGE_BYTES) >> 4
            uint32_t acc = 0;
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (!SPLIT) {
#pragma unroll
                    for (int k = 0; k < WG_PIX / 8; ++k) {  // tf32: K = 8 pixel rows per MMA (two 4-row swizzle groups)
                        const uint32_t alo = lo0 + slo + k * (1024 >> 4);
#pragma unroll
                        for (int tt = 0; tt < TPC; ++tt) {
                            const uint32_t blo =
                                lo0 + slo + ((L::OFF_B32 + tt * (BN / 32) * WG_ATOM_BYTES + k * 1024) >> 4);
                            umma_tf32_lh(tmem_base + tt * BN, alo, hi32, blo, hi32, idesc32, k == 0 ? acc : 1u);
                        }
                    }
                }
                if (SPLIT) {
#pragma unroll
                    for (int k = 0; k < WG_PIX / 16; ++k) {  // bf16: K = 16 pixel rows per MMA (two 8-row groups)
                        const uint32_t ah = lo16 + slo + ((L::OFF_A16 + k * 2048) >> 4);
                        const uint32_t al = ah + (L::A16 >> 4);
#pragma unroll
                        for (int tt = 0; tt < TPC; ++tt) {
                            const uint32_t bh = lo16 + slo + ((L::OFF_B16 + tt * (BN / 64) * WG_ATOM_BYTES + k * 2048) >> 4);
                            const uint32_t bl = bh + (L::B16 >> 4);
                            umma_bf16_lh(tmem_base + tt * BN, ah, hi16, bh, hi16, idesc16, k == 0 ? acc : 1u);  // A_h * B_h
                            umma_bf16_lh(tmem_base + tt * BN, al, hi16, bh, hi16, idesc16, 1);                  // A_l * B_h
                            umma_bf16_lh(tmem_base + tt * BN, ah, hi16, bl, hi16, idesc16, 1);                  // A_h * B_l
                        }
                    }
                }
                acc = 1;
                umma_commit(&empty_bar[stage]);
                slo += L::STAGE_BYTES >> 4;
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                    slo = 0;
                }
            }
            umma_commit(accum_bar);
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int tt = 0; tt < TPC; ++tt) {
            float* orow = p.dw + ((long long)(tap0 + tt) * p.ca + row) * p.cb + n0;
#pragma unroll 1
            for (int j = 0; j < BN / 32; ++j) {
                if (j >= b_atoms) break;
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(tt * BN + j * 32), v);
                tmem_ld_wait();
                if (row < p.ca) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        red_add_v4(orow + j * 32 + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<L::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ row-halo variant
constexpr int WR_B_ROWS = WG_PIX + 2;        // 34 haloed pixels
constexpr int WR_B_BYTES = WR_B_ROWS * 128;  // 4352
constexpr int WR_B_SLOT = 5 * 1024;          // 1024-aligned slot per 64-channel atom

struct WrParams {
    CUtensorMap amap16[2];  // dY h16 / l16: box (64 ch, 32 px, 1, 1 row, 1)
    CUtensorMap bmap16[2];  // X  h16 / l16: box (64 ch, 34 px, 1, 1 row, 1)
    int ca, cb;             // dY / X channel counts (multiples of 64)
    int n_img, gh, tiles_w; // pixel tiles: n_img * gh * tiles_w row segments of 32 pixels
    int tiles_per_cta;
    int n_tiles_b;
    float* dw;              // [9][ca][cb], accumulated with red.add (caller zeroes)
};

template <int BN, int STAGES>
struct WrSmem {
    static constexpr int A16 = 2 * WG_ATOM_BYTES;          // 128 dY channels, per plane
    static constexpr int B16 = (BN / 64) * WR_B_SLOT;      // BN X channels (haloed), per plane
    static constexpr int OFF_B = 2 * A16;
    static constexpr int STAGE_BYTES = OFF_B + 2 * B16;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
    static constexpr int TMEM_COLS = (3 * BN <= 256) ? 256 : 512;
};

// blockIdx = (split-K slice, (A tile, B tile), window row r): this CTA accumulates dW[3r + c][m0.., n0..] for c = 0..2.
template <int BN, int STAGES>
__global__ void __launch_bounds__(192) wgrad_row_kernel(const __grid_constant__ WrParams p) {
    using L = WrSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = (blockIdx.y / p.n_tiles_b) * 128;
    const int n0 = (blockIdx.y % p.n_tiles_b) * BN;
    const int dh = (int)blockIdx.z - 1;
    const int a16 = min(2, (p.ca - m0) / 64);
    const int b16 = min(BN / 64, (p.cb - n0) / 64);
    const int total_tiles = p.n_img * p.gh * p.tiles_w;
    const int tile_begin = blockIdx.x * p.tiles_per_cta;
    const int tile_end = min(total_tiles, tile_begin + p.tiles_per_cta);
    const int ntiles = tile_end - tile_begin;
    if (ntiles <= 0) return;  // uniform for the whole CTA

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<L::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t tx_bytes = 2 * (a16 * WG_ATOM_BYTES + b16 * WR_B_BYTES);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                int t = tile;
                const int w0 = (t % p.tiles_w) * WG_PIX;
                t /= p.tiles_w;
                const int h = t % p.gh;
                const int img = t / p.gh;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * L::STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                for (int pl = 0; pl < 2; ++pl) {
                    for (int a = 0; a < a16; ++a)
                        tma_load_5d(st + pl * L::A16 + a * WG_ATOM_BYTES, &p.amap16[pl], &full_bar[stage], m0 + a * 64,
                                    w0, 0, h, img);
                    for (int b = 0; b < b16; ++b)
                        tma_load_5d(st + L::OFF_B + pl * L::B16 + b * WR_B_SLOT, &p.bmap16[pl], &full_bar[stage],
                                    n0 + b * 64, w0 - 1, 0, h + dh, img);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc16 = make_idesc_bf16(128, BN, 1, 1);
            constexpr uint32_t hi16 = desc_hi_word(1024, 2);  // MN-major, 128B swizzle: 8-row groups 1024 B apart
            const uint32_t a_lo = desc_lo_word(smem_u32(smem), WG_ATOM_BYTES);            // LBO = dY atom stride
            const uint32_t b_lo = desc_lo_word(smem_u32(smem) + L::OFF_B, WR_B_SLOT);     // LBO = X slot stride
            int stage = 0;
            uint32_t phase = 0;
            uint32_t slo = 0;
            uint32_t acc = 0;
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < WG_PIX / 16; ++k) {  // K = 16 pixel rows per MMA
                    const uint32_t ah = a_lo + slo + ((k * 2048) >> 4);
                    const uint32_t al = ah + (L::A16 >> 4);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {  // window column c reads X rows [c, c + 32) of the haloed tile
                        const uint32_t bh = b_lo + slo + ((c * 128 + k * 2048) >> 4);
                        const uint32_t bl = bh + (L::B16 >> 4);
                        umma_bf16_lh(tmem_base + c * BN, ah, hi16, bh, hi16, idesc16, k == 0 ? acc : 1u);
                        umma_bf16_lh(tmem_base + c * BN, al, hi16, bh, hi16, idesc16, 1);
                        umma_bf16_lh(tmem_base + c * BN, ah, hi16, bl, hi16, idesc16, 1);
                    }
                }
                acc = 1;
                umma_commit(&empty_bar[stage]);
                slo += L::STAGE_BYTES >> 4;
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                    slo = 0;
                }
            }
            umma_commit(accum_bar);
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            float* orow = p.dw + ((long long)((dh + 1) * 3 + c) * p.ca + row) * p.cb + n0;
#pragma unroll 1
            for (int j = 0; j < BN / 32; ++j) {
                if (j >= 2 * b16) break;
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * BN + j * 32), v);
                tmem_ld_wait();
                if (row < p.ca) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        red_add_v4(orow + j * 32 + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<L::TMEM_COLS>(tmem_base);
}

}  // namespace a2x
