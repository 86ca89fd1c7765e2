// Window / grid attention of the CoBEVT fusion network on the 5th-generation tensor cores (tcgen05 + TMEM), forward and
// backward. Replaces the einsum / softmax / einsum of Attention.forward (cobevt_modules/swap_fusion_modules.py:78-127)
// and its autograd for windows of n = L*w*w tokens, n % 16 == 0, n <= 128 (config 4: 7 agents x 4x4 = 112), w = 4,
// dim_head = 32. The small windows of the V2X-ViT pyramid (4 / 16 tokens) stay on the packed SIMT kernels.
//
// A persistent CTA (128 threads) owns one head and walks over windows; thread t <-> token t <-> TMEM lane t.
// Every operand is the bf16 split pair of an fp32 row, kept side by side in ONE 128-byte shared-memory row
// [hi(32) | lo(32)] with the 128-byte swizzle, so a tile is simply 128 rows x 128 B and the UMMA descriptors decide how
// it is read:
//   * K-major  (row = M/N index, 128 B = 64 k-elements): Q and K for S = Q K^T, dO and V for dP = dO V^T; the four
//     16-element k-slices of a row are hi[0:16] hi[16:32] lo[0:16] lo[16:32], so  hi*hi + lo*hi + hi*lo  is six
//     K = 16 MMAs pairing slices (0,0) (1,1) (2,0) (3,1) (0,2) (1,3).
//   * MN-major (row = k index, 128 B = 64 n-elements): the same V / dO / Q / K tiles as the B operand of
//     O = P V, dV = P^T dO, dK = dS^T Q, dQ = dS K with N = 64: columns 0-31 accumulate (.)*hi, columns 32-63 (.)*lo,
//     and the epilogue adds the two halves.
//   * P and dS (fp32 in registers after the softmax) are written as bf16 split planes, row i = 128 B per 64 keys; read
//     K-major as the A operand of P V / dS K, and MN-major (M = key index) as the A operand of P^T dO / dS^T Q.
// S / dP live in TMEM (128 columns each), the softmax runs on the tcgen05.ld'ed fp32 row of each thread, the
// relative-position bias comes from a per-head shared-memory table, keys of padded agents are masked (-inf). The bias
// gradient needs sum over windows of dS_ij per (i, j) (the table entry of (i, j) is the same in every window): each thread
// keeps its row of that sum in 128 spare TMEM columns (tcgen05.ld / add / tcgen05.st per window, no atomics on the data
// path) and the CTA folds it into the table once at its end.
#pragma once
#include "a2x_ptx.cuh"

namespace a2x {

constexpr int WT_TILE = 128 * 128;  // bytes: 128 rows x 128 B
constexpr int WT_W = 4;             // window edge
constexpr int WT_S2 = 2 * WT_W - 1;
constexpr int WT_DH = 32;

struct WinTcParams {
    const float* qkv;      // [B*L][H][W][3*D]
    const float* dout;     // backward: [B*L][H][W][D]
    const float* bias;     // [(2L-1)(2w-1)^2][heads]
    const int* key_mask;   // [B][L] or null
    SplitOut out;          // forward: [B*L][H][W][D]; backward: dqkv [B*L][H][W][3*D]
    float* dbias;          // backward
    int B, L, H, W, heads, grid_mode;
    float scale;
};

// byte offset of 16-byte chunk c (0..7) of row r inside a 128-byte-swizzled [rows][128 B] tile
__device__ __forceinline__ uint32_t wt_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void wt_split2(float a, float b, uint32_t& h, uint32_t& l) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    h = *reinterpret_cast<const uint32_t*>(&hh);
    l = *reinterpret_cast<const uint32_t*>(&ll);
}

// 8 consecutive fp32 of a row (elements 8*c8 .. 8*c8+7 of 32) -> hi chunk c8, lo chunk 4 + c8 of tile row r
__device__ __forceinline__ void wt_store8_hl(uint8_t* tile, int r, int c8, float4 a, float4 b) {
    uint4 h, l;
    wt_split2(a.x, a.y, h.x, l.x);
    wt_split2(a.z, a.w, h.y, l.y);
    wt_split2(b.x, b.y, h.z, l.z);
    wt_split2(b.z, b.w, h.w, l.w);
    *reinterpret_cast<uint4*>(tile + wt_off(r, c8)) = h;
    *reinterpret_cast<uint4*>(tile + wt_off(r, 4 + c8)) = l;
}

// 8 consecutive probabilities of row r (keys 8*c16 .. 8*c16+7, c16 = 0..15) -> split planes (hi at `plane`, lo at plane + 2 tiles)
__device__ __forceinline__ void wt_store8_planes(uint8_t* plane, int r, int c16, const float* v) {
    uint4 h, l;
    wt_split2(v[0], v[1], h.x, l.x);
    wt_split2(v[2], v[3], h.y, l.y);
    wt_split2(v[4], v[5], h.z, l.z);
    wt_split2(v[6], v[7], h.w, l.w);
    const uint32_t off = (uint32_t)(c16 >> 3) * WT_TILE + wt_off(r, c16 & 7);
    *reinterpret_cast<uint4*>(plane + off) = h;
    *reinterpret_cast<uint4*>(plane + 2 * WT_TILE + off) = l;
}

__device__ __forceinline__ long long wt_token(const WinTcParams& p, int b, int x, int y, int X, int Y, int t) {
    const int l = t >> 4, r = t & 15, w1 = r >> 2, w2 = r & 3;
    const int ph = p.grid_mode ? w1 * X + x : x * WT_W + w1;
    const int pw = p.grid_mode ? w2 * Y + y : y * WT_W + w2;
    return ((long long)(b * p.L + l) * p.H + ph) * p.W + pw;
}

// rows of `src` (row stride `rs` floats, 32 floats used from column `c0`) -> [hi | lo] tile; optional scale
__device__ __forceinline__ void wt_load_tile(uint8_t* tile, const float* src, long long rs, int c0, const long long* sTok,
                                             int n, float scale) {
    for (int idx = threadIdx.x; idx < n * 4; idx += 128) {
        const int r = idx >> 2, c8 = idx & 3;
        const float4* g = reinterpret_cast<const float4*>(src + sTok[r] * rs + c0 + c8 * 8);
        float4 a = __ldg(g), b = __ldg(g + 1);
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
        wt_store8_hl(tile, r, c8, a, b);
    }
}

// S (or dP) = A B^T over 32 features held as [hi | lo] K-major rows: hi*hi + lo*hi + hi*lo
__device__ __forceinline__ void wt_mma_qk(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    constexpr uint32_t hi = desc_hi_word(1024, 2);
    umma_bf16_lh(tacc, a_lo + 0, hi, b_lo + 0, hi, idesc, 0);
    umma_bf16_lh(tacc, a_lo + 2, hi, b_lo + 2, hi, idesc, 1);
    umma_bf16_lh(tacc, a_lo + 4, hi, b_lo + 0, hi, idesc, 1);
    umma_bf16_lh(tacc, a_lo + 6, hi, b_lo + 2, hi, idesc, 1);
    umma_bf16_lh(tacc, a_lo + 0, hi, b_lo + 4, hi, idesc, 1);
    umma_bf16_lh(tacc, a_lo + 2, hi, b_lo + 6, hi, idesc, 1);
}

// D[128 x 64] = (Ph + Pl) [Bh | Bl]:  A = split planes of P / dS, read K-major (a_mn = 0: rows = M) or MN-major
// (a_mn = 1: rows = K); B = an [hi | lo] tile read MN-major (rows = K). nk = n / 16 k-steps.
__device__ __forceinline__ void wt_mma_pv(uint32_t tacc, uint32_t plane_addr, uint32_t b_addr, int nk, int a_mn) {
    constexpr uint32_t hi = desc_hi_word(1024, 2);
    const uint32_t idesc = make_idesc_bf16(128, 64, (uint32_t)a_mn, 1);
    const uint32_t b_lo = desc_lo_word(b_addr, WT_TILE);
    uint32_t acc = 0;
    for (int pl = 0; pl < 2; ++pl) {
        const uint32_t a_lo = desc_lo_word(plane_addr + pl * 2 * WT_TILE, a_mn ? WT_TILE : 16);
        for (int kk = 0; kk < nk; ++kk) {
            const uint32_t a_off = a_mn ? (uint32_t)kk * (2048 >> 4) : (uint32_t)(kk >> 2) * (WT_TILE >> 4) + (uint32_t)(kk & 3) * 2;
            umma_bf16_lh(tacc, a_lo + a_off, hi, b_lo + (uint32_t)kk * (2048 >> 4), hi, idesc, acc);
            acc = 1;
        }
    }
}

// the thread's score row from TMEM (+ relative-position bias, key mask) -> s[128]; returns the row maximum
__device__ __forceinline__ float wt_scores(uint32_t trow, int n, int t, uint32_t kmask, const float* sB, int L, float* s) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c * 32 < n) tmem_ld_32x32(trow + c * 32, s + c * 32);
    tmem_ld_wait();
    const int li = t >> 4, i1 = (t >> 2) & 3, i2 = t & 3;
    const int base = ((li + L - 1) * WT_S2 + (i1 + WT_W - 1)) * WT_S2 + (i2 + WT_W - 1);
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
        const int lj = j >> 4;
        const int sub = (lj * WT_S2 + ((j >> 2) & 3)) * WT_S2 + (j & 3);
        const bool ok = j < n && ((kmask >> lj) & 1u);
        const float v = ok ? s[j] + sB[base - sub] : -INFINITY;
        s[j] = v;
        m = fmaxf(m, v);
    }
    return m;
}

constexpr int WTF_SMEM = 4 * WT_TILE + WT_TILE + 4096 + 1024 + 256 + 1024;   // P planes (over Q, K) | V | bias | tokens | barrier | align

__global__ void __launch_bounds__(128, 2) window_attention_tc_fwd_kernel(const WinTcParams p, int num_windows) {
    extern __shared__ uint8_t wt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wt_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tQ = smem;                      // Q tile, later P planes (4 tiles: hi atoms 0-1, lo atoms 0-1)
    uint8_t* tK = smem + WT_TILE;
    uint8_t* tV = smem + 4 * WT_TILE;
    float* sB = reinterpret_cast<float*>(smem + 5 * WT_TILE);
    long long* sTok = reinterpret_cast<long long*>(smem + 5 * WT_TILE + 4096);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 5 * WT_TILE + 4096 + 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int n = p.L * 16;
    const int D = p.heads * WT_DH;
    const int X = p.H / WT_W, Y = p.W / WT_W;
    const int head = blockIdx.x % p.heads;
    const int G = gridDim.x / p.heads;
    const int nb = (2 * p.L - 1) * WT_S2 * WT_S2;
    for (int i = t; i < nb; i += 128) sB[i] = p.bias[i * p.heads + head];
    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)n, 0, 0);
    uint32_t ph = 0;

    for (int win = blockIdx.x / p.heads; win < num_windows; win += G) {
        const int y = win % Y, x = (win / Y) % X, b = win / (Y * X);
        if (t < n) sTok[t] = wt_token(p, b, x, y, X, Y, t);
        uint32_t kmask = 0xffffffffu;
        if (p.key_mask != nullptr) {
            kmask = 0;
            for (int l = 0; l < p.L; ++l) kmask |= (p.key_mask[b * p.L + l] != 0 ? 1u : 0u) << l;
        }
        __syncthreads();
        wt_load_tile(tQ, p.qkv, 3 * D, head * WT_DH, sTok, n, p.scale);
        wt_load_tile(tK, p.qkv, 3 * D, D + head * WT_DH, sTok, n, 1.f);
        wt_load_tile(tV, p.qkv, 3 * D, 2 * D + head * WT_DH, sTok, n, 1.f);
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            wt_mma_qk(tmem, desc_lo_word(smem_u32(tQ), 16), desc_lo_word(smem_u32(tK), 16), idesc_s);
            umma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after();
        float s[128];
        const float m = wt_scores(trow, n, t, kmask, sB, p.L, s);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
            s[j] = __expf(s[j] - m);
            sum += s[j];
        }
        // every thread has read its S row and the S MMAs have retired: P may overwrite Q / K, O may overwrite S
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c * 8 < n) wt_store8_planes(tQ, t, c, s + c * 8);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            wt_mma_pv(tmem, smem_u32(tQ), smem_u32(tV), n >> 4, 0);
            umma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after();
        float o[64];
        tmem_ld_32x32(trow, o);
        tmem_ld_32x32(trow + 32, o + 32);
        tmem_ld_wait();
        if (t < n) {
            const float inv = 1.f / sum;
            const long long off = sTok[t] * D + head * WT_DH;
#pragma unroll
            for (int c = 0; c < 32; c += 4)
                store_split4(p.out, off + c, make_float4((o[c] + o[32 + c]) * inv, (o[c + 1] + o[33 + c]) * inv,
                                                         (o[c + 2] + o[34 + c]) * inv, (o[c + 3] + o[35 + c]) * inv));
        }
        tc_fence_before();
        __syncthreads();   // tiles, token table and TMEM are reused by the next window
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// backward: tiles Q | K | V | dO, P planes (4 tiles), dS planes (4 tiles), bias, bias gradient, tokens, barrier
constexpr int WTB_SMEM = 12 * WT_TILE + 4096 + 4096 + 1024 + 256 + 1024;

__global__ void __launch_bounds__(128, 1) window_attention_tc_bwd_kernel(const WinTcParams p, int num_windows) {
    extern __shared__ uint8_t wt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wt_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tQ = smem;
    uint8_t* tK = smem + WT_TILE;
    uint8_t* tV = smem + 2 * WT_TILE;
    uint8_t* tO = smem + 3 * WT_TILE;        // dO
    uint8_t* tP = smem + 4 * WT_TILE;        // P planes
    uint8_t* tS = smem + 8 * WT_TILE;        // dS planes
    float* sB = reinterpret_cast<float*>(smem + 12 * WT_TILE);
    float* sdB = sB + 1024;
    long long* sTok = reinterpret_cast<long long*>(smem + 12 * WT_TILE + 8192);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 12 * WT_TILE + 8192 + 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int t = threadIdx.x, warp = t >> 5;
    const int n = p.L * 16;
    const int D = p.heads * WT_DH;
    const int X = p.H / WT_W, Y = p.W / WT_W;
    const int head = blockIdx.x % p.heads;
    const int G = gridDim.x / p.heads;
    const int nb = (2 * p.L - 1) * WT_S2 * WT_S2;
    for (int i = t; i < nb; i += 128) {
        sB[i] = p.bias[i * p.heads + head];
        sdB[i] = 0.f;
    }
    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t tacc = trow + 256;    // this thread's row of sum_windows dS (bias gradient), columns [256, 384)
    const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)n, 0, 0);
    const int li = t >> 4, i1 = (t >> 2) & 3, i2 = t & 3;
    const int bbase = ((li + p.L - 1) * WT_S2 + (i1 + WT_W - 1)) * WT_S2 + (i2 + WT_W - 1);
    uint32_t ph = 0;
    {
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_32x32(tacc + c * 32, z);
        tmem_st_wait();
    }

    for (int win = blockIdx.x / p.heads; win < num_windows; win += G) {
        const int y = win % Y, x = (win / Y) % X, b = win / (Y * X);
        if (t < n) sTok[t] = wt_token(p, b, x, y, X, Y, t);
        uint32_t kmask = 0xffffffffu;
        if (p.key_mask != nullptr) {
            kmask = 0;
            for (int l = 0; l < p.L; ++l) kmask |= (p.key_mask[b * p.L + l] != 0 ? 1u : 0u) << l;
        }
        __syncthreads();
        wt_load_tile(tQ, p.qkv, 3 * D, head * WT_DH, sTok, n, p.scale);
        wt_load_tile(tK, p.qkv, 3 * D, D + head * WT_DH, sTok, n, 1.f);
        wt_load_tile(tV, p.qkv, 3 * D, 2 * D + head * WT_DH, sTok, n, 1.f);
        wt_load_tile(tO, p.dout, D, head * WT_DH, sTok, n, 1.f);
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            wt_mma_qk(tmem, desc_lo_word(smem_u32(tQ), 16), desc_lo_word(smem_u32(tK), 16), idesc_s);          // S
            wt_mma_qk(tmem + 128, desc_lo_word(smem_u32(tO), 16), desc_lo_word(smem_u32(tV), 16), idesc_s);    // dP = dO V^T
            umma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after();
        float s[128];
        const float m = wt_scores(trow, n, t, kmask, sB, p.L, s);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
            s[j] = __expf(s[j] - m);
            sum += s[j];
        }
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 128; ++j) s[j] *= inv;                       // P
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c * 8 < n) wt_store8_planes(tP, t, c, s + c * 8);
        float Dv = 0.f;                                                   // D_i = sum_j P_ij dP_ij
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 32 < n) {
                float dp[32];
                tmem_ld_32x32(trow + 128 + c * 32, dp);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c * 32 + j < n) Dv = fmaf(s[c * 32 + j], dp[j], Dv);        // TMEM columns >= n are stale (may be NaN)
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c * 32 < n) {
                float dp[32], ac[32];
                tmem_ld_32x32(trow + 128 + c * 32, dp);
                tmem_ld_32x32(tacc + c * 32, ac);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float ds = c * 32 + j < n ? s[c * 32 + j] * (dp[j] - Dv) : 0.f;   // dS_ij (0 at masked keys)
                    dp[j] = ds;
                    ac[j] += ds;
                }
                tmem_st_32x32(tacc + c * 32, ac);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) wt_store8_planes(tS, t, c * 4 + c8, dp + c8 * 8);
            }
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            const int nk = n >> 4;
            wt_mma_pv(tmem, smem_u32(tP), smem_u32(tO), nk, 1);           // dV = P^T dO
            wt_mma_pv(tmem + 64, smem_u32(tS), smem_u32(tQ), nk, 1);      // dK = dS^T (scale Q)
            wt_mma_pv(tmem + 128, smem_u32(tS), smem_u32(tK), nk, 0);     // dQ = dS K (scaled below)
            umma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int part = 0; part < 3; ++part) {                            // 0: dV, 1: dK, 2: dQ
            float o[64];
            tmem_ld_32x32(trow + part * 64, o);
            tmem_ld_32x32(trow + part * 64 + 32, o + 32);
            tmem_ld_wait();
            if (t < n) {
                const float f = part == 2 ? p.scale : 1.f;
                const long long off = sTok[t] * (3 * D) + (2 - part) * D + head * WT_DH;
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    store_split4(p.out, off + c, make_float4((o[c] + o[32 + c]) * f, (o[c + 1] + o[33 + c]) * f,
                                                             (o[c + 2] + o[34 + c]) * f, (o[c + 3] + o[35 + c]) * f));
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    __syncthreads();
    // fold the per-thread rows of sum dS into the head's table (entry of (i, j) = bbase_i - sub_j), then one global
    // atomic per entry and CTA
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c * 32 < n) {
            float ac[32];
            tmem_ld_32x32(tacc + c * 32, ac);
            tmem_ld_wait();
            if (t < n) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int jj = c * 32 + j;
                    if (jj < n && ac[j] != 0.f)
                        atomicAdd(&sdB[bbase - (((jj >> 4) * WT_S2 + ((jj >> 2) & 3)) * WT_S2 + (jj & 3))], ac[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    for (int i = t; i < nb; i += 128)
        if (sdB[i] != 0.f) atomicAdd(&p.dbias[i * p.heads + head], sdB[i]);
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace a2x
