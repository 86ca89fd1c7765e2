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
#include "philox.cuh"

namespace a2x {

constexpr int TG_BM = 128;  // pixels per tile (UMMA M)
constexpr int TG_A_BYTES = TG_BM * 128;
constexpr int TG_MAX_TAPS = 27;  // 9 window taps x 3 split products
constexpr int TG_MAX_MAPS = 12;  // 4 parity views x {hi fp32, h16, l16}
// Epilogue stores go through a per-warp shared-memory transpose: a thread owns one pixel ROW of the accumulator, so direct
// stores would put 32 lanes on 32 different 128-byte lines, 16 bytes each (partial sectors: measured 3-4x slower than the
// MMAs of a 1x1 / stride-2 tile). The transposed mapping gives every store instruction full 32-byte sectors: 4 lanes x 16 B
// of one pixel's 16 fp32 columns, 2 lanes x 16 B of its 16 bf16 columns.
constexpr int TG_XP_STRIDE = 20;                       // floats per row of the [32 pixels][16 columns] tile (+4 pad: conflict-free)
constexpr int TG_XP_BYTES = 32 * TG_XP_STRIDE * 4;     // 2560 B per epilogue warp

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
    int relu;        // activation: 0 none, 1 ReLU, 2 GELU (erf)
    int accumulate;  // out += result (residual add; the fp32 plane holds the previous value)
    // fused BatchNorm batch statistics of the raw result: stats[ch] += sum, stats[stat_c + ch] += sum of squares
    double* stats;
    int stat_c;
    // x + dropout(linear(y) + bias) of the transformer sublayers in ONE epilogue (fp32-only outputs): `res` is the residual
    // (same addressing as the output, may be a different buffer), drop.thresh != 0 turns the Philox mask on
    const float* res;
    DropArgs drop;
    long long drop_elem0;   // mask element index of the output's element 0 (the output may be a slice of the site's tensor)
};

constexpr int TG_EPW = 8;   // epilogue warps of tapgemm_kernel: two per TMEM lane quadrant, alternating 32-column chunks

template <int BN, int STAGES>
struct TgSmem {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = TG_A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int NBAR = 2 * STAGES + 4;  // full/empty per stage + tmem_full[2] + tmem_empty[2]
    static constexpr int STAT_OFF = BAR_OFF + NBAR * 8 + 16;
    static constexpr int XP_OFF = STAT_OFF + 4 * 2 * BN * 4;         // [4 warps][2][BN] stats, then the store-transpose tiles
    static constexpr int TOTAL = XP_OFF + TG_EPW * TG_XP_BYTES + 1024;    // + alignment slack
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                     : (2 * BN <= 256) ? 256 : 512;  // two accumulator buffers
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

// Common epilogue arguments (tapgemm and its halo variant share the drain code)
struct TgEpi {
    SplitOut out;
    long long osn, osh, osw;
    int sub_c, sub_s;
    long long sub_sh, sub_sw;
    int ncols;
    const float* scale;
    const float* shift;
    int relu, accumulate;
    double* stats;
    int stat_c;
    int gh, gw;
    // BN(train)+ReLU backward reduction fused into a data-gradient epilogue: when bz != null the statistics of
    // g = v * [bz*bscale+bshift > 0] and g * zhat (zhat = (bz - bmean) * binvstd) replace the plain sum / sum of squares,
    // i.e. stats[c] += sum g, stats[stat_c + c] += sum g*zhat: pass 1 of the consumer's bn_relu_bwd for free.
    const float* bz;        // the consumer layer's pre-BN activation, same pixel grid / strides as the output
    const float* bscale;
    const float* bshift;
    const float* bmean;
    const float* binvstd;
    const float* res;       // residual read from here instead of the output's own fp32 plane (fp32-only outputs)
    DropArgs drop;          // thresh != 0: v *= keep / (1 - p) before the residual add
    long long drop_elem0;   // mask element index = drop_elem0 + output offset
};

// Drain one 128 x BN accumulator (TMEM buffer at `tacc`) for the tile at (img, h0, w0), column offset n0.
// Called by the EPW epilogue warps (EPW = 4: one per TMEM lane quadrant `q` = warp % 4; EPW = 8: two per quadrant, warp
// half `qh` takes the 32-column chunks j = qh, qh + 2, ...). `et` = thread index among the epilogue threads.
template <int BN, int EPW = 4>
__device__ __forceinline__ void tg_epilogue(const TgEpi& e, uint32_t tacc, int q, int lane, int et, int img, int h0,
                                            int w0, int tw_log2, int n0, float* sstat, float* sT, int qh = 0) {
    const int row = q * 32 + lane;
    const int h = h0 + (row >> tw_log2);
    const int w = w0 + (row & ((1 << tw_log2) - 1));
    const bool valid = (h < e.gh) && (w < e.gw);
    const long long obase = (long long)img * e.osn + (long long)h * e.osh + (long long)w * e.osw;
    // transposed store mappings (see TG_XP_STRIDE): the pixels this lane stores are fixed per tile
    long long offA[4], offB[2];
    unsigned validA = 0, validB = 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int prow = q * 32 + 8 * jj + (lane >> 2);
        const int ph = h0 + (prow >> tw_log2), pw = w0 + (prow & ((1 << tw_log2) - 1));
        validA |= (ph < e.gh && pw < e.gw ? 1u : 0u) << jj;
        offA[jj] = (long long)img * e.osn + (long long)ph * e.osh + (long long)pw * e.osw + 4 * (lane & 3);
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int prow = q * 32 + 16 * jj + (lane >> 1);
        const int ph = h0 + (prow >> tw_log2), pw = w0 + (prow & ((1 << tw_log2) - 1));
        validB |= (ph < e.gh && pw < e.gw ? 1u : 0u) << jj;
        offB[jj] = (long long)img * e.osn + (long long)ph * e.osh + (long long)pw * e.osw + 8 * (lane & 1);
    }
#pragma unroll 1
    for (int j = qh; j < BN / 32; j += EPW / 4) {
        const int c0 = n0 + j * 32;
        if (c0 >= e.ncols) break;
        float v[32];
        tmem_ld_32x32(tacc + (uint32_t(q * 32) << 16) + uint32_t(j * 32), v);
        tmem_ld_wait();
        const int sub = c0 / e.sub_c;
        const int cc = c0 - sub * e.sub_c;
        if (e.scale != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= __ldg(e.scale + cc + i);
        }
        if (e.shift != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(e.shift + cc + i);
        }
        const bool do_store = e.relu != 77;  // relu == 77: debug switch, skip the stores (epilogue cost experiment)
        const long long sub_off = (long long)(sub / e.sub_s) * e.sub_sh + (long long)(sub % e.sub_s) * e.sub_sw + cc;
        if (e.out.b16 == nullptr) {
            // fp32 output only (data gradients): a thread's eight 16-byte stores cover one 128-byte line back to back and are
            // merged on the way out; the direct row stores measured faster than the transpose here
            if (valid && do_store) {
                const long long off = obase + sub_off;
                if (e.drop.thresh != 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float m[8];
                        drop_mult8(e.drop, ((e.drop_elem0 + off) >> 3) + g, m);
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[8 * g + k] *= m[k];
                    }
                }
                if (e.accumulate || e.res != nullptr) {
                    const float4* o4 = reinterpret_cast<const float4*>((e.res != nullptr ? e.res : e.out.hi) + off);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 prev = o4[i];
                        v[4 * i] += prev.x;
                        v[4 * i + 1] += prev.y;
                        v[4 * i + 2] += prev.z;
                        v[4 * i + 3] += prev.w;
                    }
                }
                if (e.relu == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (e.relu == 2) {  // exact (erf) GELU: nn.GELU() of the transformer feed-forward layers
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752f));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(e.out.hi + off + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        } else {
            // activation first when there is no residual (thread-per-row domain); with a residual it follows the add below
            if (!e.accumulate) {
                if (e.relu == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (e.relu == 2) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752f));
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(sT + lane * TG_XP_STRIDE + 4 * i) =
                        make_float4(v[16 * half + 4 * i], v[16 * half + 4 * i + 1], v[16 * half + 4 * i + 2], v[16 * half + 4 * i + 3]);
                __syncwarp();
                if (e.out.hi != nullptr) {
                    // fp32 plane: lane -> (pixel 8 jj + lane / 4, columns 4 (lane % 4) .. + 3): 64 contiguous bytes per pixel
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int p = 8 * jj + (lane >> 2), f4 = lane & 3;
                        const bool pvalid = (validA >> jj) & 1u;
                        const long long poff = offA[jj] + sub_off + 16 * half;
                        float4 x = *reinterpret_cast<const float4*>(sT + p * TG_XP_STRIDE + 4 * f4);
                        if (e.accumulate) {
                            if (pvalid) {
                                const float4 prev = *reinterpret_cast<const float4*>(e.out.hi + poff);
                                x.x += prev.x; x.y += prev.y; x.z += prev.z; x.w += prev.w;
                            }
                            if (e.relu == 1) {
                                x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                            } else if (e.relu == 2) {
                                x.x = 0.5f * x.x * (1.f + erff(x.x * 0.70710678118654752f));
                                x.y = 0.5f * x.y * (1.f + erff(x.y * 0.70710678118654752f));
                                x.z = 0.5f * x.z * (1.f + erff(x.z * 0.70710678118654752f));
                                x.w = 0.5f * x.w * (1.f + erff(x.w * 0.70710678118654752f));
                            }
                            *reinterpret_cast<float4*>(sT + p * TG_XP_STRIDE + 4 * f4) = x;   // the planes below split the sum
                        }
                        if (pvalid && do_store) *reinterpret_cast<float4*>(e.out.hi + poff) = x;
                    }
                    if (e.accumulate) __syncwarp();
                }
                // bf16 planes: lane -> (pixel 16 jj + lane / 2, columns 8 (lane % 2) .. + 7): one full 32-byte sector per pixel
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int p = 16 * jj + (lane >> 1), h8 = lane & 1;
                    const bool pvalid = (validB >> jj) & 1u;
                    const long long poff = offB[jj] + sub_off + 16 * half;
                    const float4 a = *reinterpret_cast<const float4*>(sT + p * TG_XP_STRIDE + 8 * h8);
                    const float4 b = *reinterpret_cast<const float4*>(sT + p * TG_XP_STRIDE + 8 * h8 + 4);
                    __nv_bfloat162 h01 = __floats2bfloat162_rn(a.x, a.y), h23 = __floats2bfloat162_rn(a.z, a.w);
                    __nv_bfloat162 h45 = __floats2bfloat162_rn(b.x, b.y), h67 = __floats2bfloat162_rn(b.z, b.w);
                    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                    const float2 f45 = __bfloat1622float2(h45), f67 = __bfloat1622float2(h67);
                    __nv_bfloat162 l01 = __floats2bfloat162_rn(a.x - f01.x, a.y - f01.y), l23 = __floats2bfloat162_rn(a.z - f23.x, a.w - f23.y);
                    __nv_bfloat162 l45 = __floats2bfloat162_rn(b.x - f45.x, b.y - f45.y), l67 = __floats2bfloat162_rn(b.z - f67.x, b.w - f67.y);
                    uint4 hp, lp;
                    hp.x = *reinterpret_cast<uint32_t*>(&h01); hp.y = *reinterpret_cast<uint32_t*>(&h23);
                    hp.z = *reinterpret_cast<uint32_t*>(&h45); hp.w = *reinterpret_cast<uint32_t*>(&h67);
                    lp.x = *reinterpret_cast<uint32_t*>(&l01); lp.y = *reinterpret_cast<uint32_t*>(&l23);
                    lp.z = *reinterpret_cast<uint32_t*>(&l45); lp.w = *reinterpret_cast<uint32_t*>(&l67);
                    if (pvalid && do_store) {
                        *reinterpret_cast<uint4*>(e.out.b16 + poff) = hp;
                        *reinterpret_cast<uint4*>(e.out.b16 + e.out.ps + poff) = lp;
                    }
                }
                __syncwarp();   // the tile is rewritten by the next half / chunk
            }
        }
        if (e.stats != nullptr) {  // warp-uniform
            float sq[32];
            if (e.bz != nullptr) {
                const long long off = obase + cc;  // plain (non-scattered) output addressing
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) z4 = *reinterpret_cast<const float4*>(e.bz + off + 4 * i);
                    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ch = cc + 4 * i + k;
                        const float z = zz[k];
                        const float g = (valid && z * __ldg(e.bscale + ch) + __ldg(e.bshift + ch) > 0.f) ? v[4 * i + k] : 0.f;
                        v[4 * i + k] = g;
                        sq[4 * i + k] = g * (z - __ldg(e.bmean + ch)) * __ldg(e.binvstd + ch);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = valid ? v[i] : 0.f;
                    sq[i] = v[i] * v[i];
                }
            }
            const float s1 = warp_transpose_sum(v, lane);
            const float s2 = warp_transpose_sum(sq, lane);
            sstat[(q * 2 + 0) * BN + j * 32 + lane] = s1;  // each (warp, column) written exactly once:
            sstat[(q * 2 + 1) * BN + j * 32 + lane] = s2;  // fixed-order combine below => deterministic
        }
    }
    if (e.stats != nullptr) {
        asm volatile("bar.sync 1, %0;" ::"n"(EPW * 32) : "memory");
        for (int i = et; i < BN; i += EPW * 32) {
            const int col = n0 + i;
            if (col < e.ncols) {
                const int ch = col % e.sub_c;
                const float t1 = (sstat[i] + sstat[2 * BN + i]) + (sstat[4 * BN + i] + sstat[6 * BN + i]);
                const float t2 = (sstat[BN + i] + sstat[3 * BN + i]) + (sstat[5 * BN + i] + sstat[7 * BN + i]);
                atomicAdd(&e.stats[ch], (double)t1);
                atomicAdd(&e.stats[e.stat_c + ch], (double)t2);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPW * 32) : "memory");  // sstat is reused by the next tile
    }
}

// Persistent: gridDim.x CTAs loop over (pixel tile, column tile) pairs; the smem ring runs across tiles and the
// TMEM accumulator is double buffered, so the epilogue of tile i overlaps the MMAs of tile i + 1.
template <int BN, int STAGES>
__global__ void __launch_bounds__(64 + TG_EPW * 32) tapgemm_kernel(const __grid_constant__ TgParams p) {
    using L = TgSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* sstat = reinterpret_cast<float*>(smem + L::STAT_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int TW = 1 << p.tw_log2;
    const int TH = TG_BM >> p.tw_log2;
    const int tiles_m = p.n_img * p.tiles_h * p.tiles_w;
    const int tiles_n = (p.ncols + BN - 1) / BN;
    const int n_tiles = tiles_m * tiles_n;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], TG_EPW);  // one arrival per epilogue warp
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
            tma_prefetch_desc(&p.bmap);
            tma_prefetch_desc(&p.amap[0]);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int t = tile / tiles_n;                      // column tile fastest: the CTAs that share a pixel tile run together,
                const int n0 = (tile % tiles_n) * BN;        // so its A operand is fetched from HBM once (wide linears: 3 - 5 column tiles)
                const int tw_i = t % p.tiles_w;
                t /= p.tiles_w;
                const int th_i = t % p.tiles_h;
                const int img = t / p.tiles_h;
                const int h0 = th_i * TH, w0 = tw_i * TW;
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
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc32 = make_idesc_tf32(TG_BM, BN, 0, 0);
            constexpr uint32_t idesc16 = make_idesc_bf16(TG_BM, BN, 0, 0);
            constexpr uint32_t dhi = desc_hi_word(1024, 2);  // K-major SW128: SBO = 1024, layout 2
            const uint32_t a_lo0 = desc_lo_word(smem_u32(smem), 16);
            const uint32_t b_lo0 = desc_lo_word(smem_u32(smem) + TG_A_BYTES, 16);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t slo = 0;  // (stage * STAGE_BYTES) >> 4, carried incrementally
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t tacc = tmem_base + buf * BN;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);  // epilogue drained this buffer
                tc_fence_after();
                uint32_t acc = 0;
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    const int kind = p.taps[tap].kind;
                    const int nk = kind ? p.kchunks16 : p.kchunks32;
                    for (int kc = 0; kc < nk; ++kc) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t alo = a_lo0 + slo, blo = b_lo0 + slo;
                        if (kind) {
                            umma_bf16_lh(tacc, alo, dhi, blo, dhi, idesc16, acc);
                            umma_bf16_lh(tacc, alo + 2, dhi, blo + 2, dhi, idesc16, 1);
                            umma_bf16_lh(tacc, alo + 4, dhi, blo + 4, dhi, idesc16, 1);
                            umma_bf16_lh(tacc, alo + 6, dhi, blo + 6, dhi, idesc16, 1);
                        } else {
                            umma_tf32_lh(tacc, alo, dhi, blo, dhi, idesc32, acc);
                            umma_tf32_lh(tacc, alo + 2, dhi, blo + 2, dhi, idesc32, 1);
                            umma_tf32_lh(tacc, alo + 4, dhi, blo + 4, dhi, idesc32, 1);
                            umma_tf32_lh(tacc, alo + 6, dhi, blo + 6, dhi, idesc32, 1);
                        }
                        acc = 1;
                        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                        slo += L::STAGE_BYTES >> 4;
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                            slo = 0;
                        }
                    }
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        // epilogue warps 2..9 -> TMEM lane quadrant (warp % 4), column-chunk half (warp - 2) / 4
        const int q = warp & 3;
        const int qh = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;
        TgEpi e;
        e.out = p.out; e.osn = p.osn; e.osh = p.osh; e.osw = p.osw; e.sub_c = p.sub_c; e.sub_s = p.sub_s;
        e.sub_sh = p.sub_sh; e.sub_sw = p.sub_sw; e.ncols = p.ncols; e.scale = p.scale; e.shift = p.shift;
        e.relu = p.relu; e.accumulate = p.accumulate; e.stats = p.stats; e.stat_c = p.stat_c; e.gh = p.gh; e.gw = p.gw;
        e.res = p.res; e.drop = p.drop; e.drop_elem0 = p.drop_elem0;
        e.bz = nullptr; e.bscale = e.bshift = e.bmean = e.binvstd = nullptr;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            int t = tile / tiles_n;
            const int n0 = (tile % tiles_n) * BN;
            const int tw_i = t % p.tiles_w;
            t /= p.tiles_w;
            const int th_i = t % p.tiles_h;
            const int img = t / p.tiles_h;
            const int buf = it & 1;
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            tg_epilogue<BN, TG_EPW>(e, tmem_base + buf * BN, q, lane, et, img, th_i * TH, tw_i * TW, p.tw_log2, n0, sstat,
                                    reinterpret_cast<float*>(smem + L::XP_OFF + (warp - 2) * TG_XP_BYTES), qh);
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
