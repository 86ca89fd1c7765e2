// Window / grid attention of the CoBEVT fusion network on the 5th-generation tensor cores (tcgen05 + TMEM), forward and
// backward. Replaces the einsum / softmax / einsum of Attention.forward (cobevt_modules/swap_fusion_modules.py:78-127)
// and its autograd for windows of n = L*w*w tokens, n % 16 == 0, n <= 128 (config 4: 7 agents x 4x4 = 112), w = 4,
// dim_head = 32. The small windows of the V2X-ViT pyramid (4 / 16 tokens) stay on the packed SIMT kernels.
//
// A persistent CTA owns one head and walks over windows. Two warpgroups: the LOADER (warps 4-7) gathers the q / k / v (/ dO)
// rows of the next window from global memory, splits them and fills one of two tile buffers; the COMPUTE group (warps
// 0-7: WT_NS = 2 threads r, r + 128 share token r = TMEM lane r; thread h takes the 16-column key groups (one agent's 4 x 4
// tokens) g = 2 k + h and 16 output features, exchanging the row max / sum / D through shared memory) issues the MMAs,
// runs the softmax and the epilogue. Keys beyond the last valid agent (the padding of the agent axis) are never touched:
// the score / probability / gradient tiles have ne = 16 (last valid agent + 1) columns, K / V rows beyond ne are not loaded,
// and the interleaved group ownership splits the remaining agents evenly between the two threads of a row. full / empty mbarriers
// per buffer (empty is arrived by tcgen05.commit of the window's last MMA), so global latency is off the compute path.
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

constexpr int WT_W = 4;             // window edge
constexpr int WT_S2 = 2 * WT_W - 1;
constexpr int WT_DH = 32;
constexpr int WT_NS = 2;                    // compute threads per score row (4 measured no faster: the chain is sync / MMA latency)
constexpr int WT_CPT = 128 / WT_NS;         // score columns per thread
constexpr int WT_GPT = WT_CPT / 16;         // 16-column key groups (one agent's 4 x 4 tokens) per thread: share h owns the
                                            // groups g = WT_NS k + h, so that the valid agents (a prefix) split evenly between
                                            // the threads of a row; register slot 16 k + i <-> key column 16 (WT_NS k + h) + i
constexpr int WT_DPT = WT_DH / WT_NS;       // output features per thread
constexpr int WT_THREADS = 128 * (WT_NS + 1);
constexpr int WT_AUX = 4096 + 4096 + 2 * 1024 + 256 + 3 * WT_NS * 512;   // bias | bias gradient | token tables (2) | barriers, TMEM slot | row max / sum / D exchange

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

// A tile is n rows x 128 B (TS = n * 128 bytes, a multiple of 1024) with the 128-byte swizzle.
// byte offset of 16-byte chunk c (0..7) of row r:
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

// 8 consecutive probabilities of row r (keys 8*c16 .. 8*c16+7, c16 = 0..15) -> split planes: hi atoms at plane + {0, TS},
// lo atoms at plane + {2 TS, 3 TS}
__device__ __forceinline__ void wt_store8_planes(uint8_t* plane, uint32_t TS, int r, int c16, const float* v) {
    uint4 h, l;
    wt_split2(v[0], v[1], h.x, l.x);
    wt_split2(v[2], v[3], h.y, l.y);
    wt_split2(v[4], v[5], h.z, l.z);
    wt_split2(v[6], v[7], h.w, l.w);
    const uint32_t off = (uint32_t)(c16 >> 3) * TS + wt_off(r, c16 & 7);
    *reinterpret_cast<uint4*>(plane + off) = h;
    *reinterpret_cast<uint4*>(plane + 2 * TS + off) = l;
}

__device__ __forceinline__ long long wt_token(const WinTcParams& p, int b, int x, int y, int X, int Y, int t) {
    const int l = t >> 4, r = t & 15, w1 = r >> 2, w2 = r & 3;
    const int ph = p.grid_mode ? w1 * X + x : x * WT_W + w1;
    const int pw = p.grid_mode ? w2 * Y + y : y * WT_W + w2;
    return ((long long)(b * p.L + l) * p.H + ph) * p.W + pw;
}

__device__ __forceinline__ void wt_bar_load() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
__device__ __forceinline__ void wt_bar() { asm volatile("bar.sync 1, %0;" ::"n"(128 * WT_NS) : "memory"); }

// NT tensors' rows (32 floats each, tensor k at `src[k]` + token * rs[k]) -> [hi | lo] tiles at tile0 + toff[k].
// ALL global loads of the window are issued before the first conversion (n * 4 <= 512 items of 32 B per tensor for the
// 128 loader threads: one round trip to L2 / HBM per window, 2 * 4 * NT float4 registers in flight).
template <int NT>
__device__ __forceinline__ void wt_load_tiles(uint8_t* tile0, const uint32_t (&toff)[NT], const float* const (&src)[NT],
                                              const long long (&rs)[NT], const float (&scale)[NT],
                                              const long long* sTok, const int (&nrows)[NT], int tl) {
#pragma unroll
    for (int k0 = 0; k0 < NT; k0 += 2) {      // two tensors (16 float4 registers) in flight at a time
        float4 a[2][4], b[2][4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = tl + it * 128;
            {
                const long long tok = sTok[idx >> 2];      // all 128 slots of the table hold a valid token index
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k0 + k < NT && idx < nrows[k0 + k] * 4) {
                        const float4* g = reinterpret_cast<const float4*>(src[k0 + k] + tok * rs[k0 + k] + (idx & 3) * 8);
                        a[k][it] = __ldg(g);
                        b[k][it] = __ldg(g + 1);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k0 + k < NT) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int idx = tl + it * 128;
                    if (idx < nrows[k0 + k] * 4) {
                        float4 x = a[k][it], y = b[k][it];
                        const float sc = scale[k0 + k];
                        x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
                        y.x *= sc; y.y *= sc; y.z *= sc; y.w *= sc;
                        wt_store8_hl(tile0 + toff[k0 + k], idx >> 2, idx & 3, x, y);
                    }
                }
            }
        }
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
__device__ __forceinline__ void wt_mma_pv(uint32_t tacc, uint32_t plane_addr, uint32_t b_addr, uint32_t TS, int nk, int a_mn) {
    constexpr uint32_t hi = desc_hi_word(1024, 2);
    const uint32_t idesc = make_idesc_bf16(128, 64, (uint32_t)a_mn, 1);
    const uint32_t b_lo = desc_lo_word(b_addr, TS);
    uint32_t acc = 0;
    for (int pl = 0; pl < 2; ++pl) {
        const uint32_t a_lo = desc_lo_word(plane_addr + pl * 2 * TS, a_mn ? TS : 16);
        for (int kk = 0; kk < nk; ++kk) {
            const uint32_t a_off = a_mn ? (uint32_t)kk * (2048 >> 4) : (uint32_t)(kk >> 2) * (TS >> 4) + (uint32_t)(kk & 3) * 2;
            umma_bf16_lh(tacc, a_lo + a_off, hi, b_lo + (uint32_t)kk * (2048 >> 4), hi, idesc, acc);
            acc = 1;
        }
    }
}

// this thread's share (key groups g = WT_NS k + h < n / 16) of score row r from TMEM (+ relative-position bias) -> s[CPT];
// groups of masked agents and groups beyond n get -inf without touching TMEM or the table. Returns the maximum over the
// share; `gv` = bit k set when group k holds keys.
__device__ __forceinline__ float wt_scores(uint32_t trow, int n, int r, int h, uint32_t kmask, const float* sB, int L, float* s,
                                           uint32_t& gv) {
    gv = 0;
#pragma unroll
    for (int k = 0; k < WT_GPT; ++k) {
        const int g = WT_NS * k + h;
        if (16 * g < n && ((kmask >> g) & 1u)) {
            tmem_ld_32x16(trow + 16 * g, s + 16 * k);
            gv |= 1u << k;
        }
    }
    tmem_ld_wait();
    const int li = r >> 4, i1 = (r >> 2) & 3, i2 = r & 3;
    const int base = ((li + L - 1) * WT_S2 + (i1 + WT_W - 1)) * WT_S2 + (i2 + WT_W - 1) - h * WT_S2 * WT_S2;
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < WT_GPT; ++k) {
        if ((gv >> k) & 1u) {      // uniform over the CTA's threads of this share
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int sub = ((WT_NS * k) * WT_S2 + ((i >> 2) & 3)) * WT_S2 + (i & 3);
                const float v = s[16 * k + i] + sB[base - sub];
                s[16 * k + i] = v;
                m = fmaxf(m, v);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[16 * k + i] = -INFINITY;
        }
    }
    return m;
}

__device__ __forceinline__ void wt_tmem_ld_dpt(uint32_t taddr, float* v) {
    if constexpr (WT_DPT == 16) tmem_ld_32x16(taddr, v);
    else tmem_ld_32x8(taddr, v);
}

// keys beyond the last valid agent are never touched (padded agents trail: CoBEVT's key mask is the padding of the agent
// axis): the score / probability / gradient tiles have ne = 16 * (highest valid agent + 1) columns instead of n
__device__ __forceinline__ int wt_keys(uint32_t kmask, int L) {
    const uint32_t m = kmask & (L >= 32 ? 0xffffffffu : ((1u << L) - 1u));
    return m == 0 ? 16 : 16 * (32 - __clz(m));
}
__device__ __forceinline__ uint32_t wt_kmask(const WinTcParams& p, int b) {
    if (p.key_mask == nullptr) return 0xffffffffu;
    uint32_t k = 0;
    for (int l = 0; l < p.L; ++l) k |= (p.key_mask[b * p.L + l] != 0 ? 1u : 0u) << l;
    return k;
}

// shared memory: forward 2 buffers x [Q | K | . | . | V] (P planes overlay Q, K and the two spare tiles) = 10 TS;
// backward 2 buffers x [Q | K | V | dO] + ONE set of split planes (4 TS) that holds P for dV = P^T dO and then dS for
// dK / dQ = 12 TS (168 KB at 112 tokens, 192 KB at 128); + WT_AUX + 1 KB alignment slack
__host__ __device__ constexpr int wt_smem_bytes(int n, bool bwd) { return (bwd ? 12 : 10) * n * 128 + WT_AUX + 1024; }

struct WtSmem {
    uint8_t* base;
    float* sB;
    float* sdB;
    long long* sTok;      // [2][128]
    uint64_t *full, *empty, *mma;
    uint32_t* tmem_slot;
    float* sEx;           // [3][NS][128]: row max | row sum | D, one value per (column share, row)
};
__device__ __forceinline__ WtSmem wt_carve(uint8_t* raw, int tiles_bytes) {
    WtSmem w;
    w.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* aux = w.base + tiles_bytes;
    w.sB = reinterpret_cast<float*>(aux);
    w.sdB = reinterpret_cast<float*>(aux + 4096);
    w.sTok = reinterpret_cast<long long*>(aux + 8192);
    w.full = reinterpret_cast<uint64_t*>(aux + 8192 + 2048);
    w.empty = w.full + 2;
    w.mma = w.empty + 2;
    w.tmem_slot = reinterpret_cast<uint32_t*>(w.mma + 1);
    w.sEx = reinterpret_cast<float*>(aux + 8192 + 2048 + 256);
    return w;
}

template <bool BWD>
__global__ void __launch_bounds__(WT_THREADS, 1) window_attention_tc_kernel(const WinTcParams p, int num_windows) {
    extern __shared__ uint8_t wt_raw[];
    const int n = p.L * 16;
    const uint32_t TS = (uint32_t)n * 128;
    constexpr int BUF_TILES = BWD ? 4 : 5;
    const WtSmem sm = wt_carve(wt_raw, (BWD ? 12 : 10) * (int)TS);
    uint8_t* tP = sm.base + 8 * TS;       // backward: the split planes (P, then dS)
    const int tid = threadIdx.x, warp = tid >> 5;
    const int D = p.heads * WT_DH;
    const int X = p.H / WT_W, Y = p.W / WT_W;
    const int head = blockIdx.x % p.heads;
    const int G = gridDim.x / p.heads;
    const int nb = (2 * p.L - 1) * WT_S2 * WT_S2;
    for (int i = tid; i < nb; i += WT_THREADS) {
        sm.sB[i] = p.bias[i * p.heads + head];
        sm.sdB[i] = 0.f;
    }
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.full[b], 128);
            mbar_init(&sm.empty[b], 1);
        }
        mbar_init(sm.mma, 1);
        fence_mbar_init();
    }
    constexpr uint32_t TM_COLS = BWD ? 512 : 128;
    if (warp == 0) tmem_alloc<TM_COLS>(sm.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_slot;

    if (warp >= 4 * WT_NS) {
        // ------------------------------------------------------------------ loader warpgroup
        const int tl = tid - 128 * WT_NS;
        int it = 0;
        int lb = -1, lne = n;      // scene of the previous window, its effective key count
        for (int win = blockIdx.x / p.heads; win < num_windows; win += G, ++it) {
            const int buf = it & 1;
            uint8_t* tb = sm.base + buf * BUF_TILES * TS;
            const int y = win % Y, x = (win / Y) % X, b = win / (Y * X);
            mbar_wait(&sm.empty[buf], ((it >> 1) & 1) ^ 1);     // the MMAs that read this buffer have retired
            sm.sTok[buf * 128 + tl] = wt_token(p, b, x, y, X, Y, tl < n ? tl : 0);
            if (b != lb) {
                lne = wt_keys(wt_kmask(p, b), p.L);
                lb = b;
            }
            wt_bar_load();
            const float* q0 = p.qkv + head * WT_DH;
            if (BWD) {
                const uint32_t toff[4] = {0, TS, 2 * TS, 3 * TS};
                const float* const src[4] = {q0, q0 + D, q0 + 2 * D, p.dout + head * WT_DH};
                const long long rs[4] = {3LL * D, 3LL * D, 3LL * D, (long long)D};
                const float sc[4] = {p.scale, 1.f, 1.f, 1.f};
                const int nr[4] = {n, lne, lne, n};   // K / V rows of the trailing masked agents are not needed
                wt_load_tiles<4>(tb, toff, src, rs, sc, sm.sTok + buf * 128, nr, tl);   // loads of 2 tensors in flight at a time
            } else {
                const uint32_t toff[3] = {0, TS, 4 * TS};
                const float* const src[3] = {q0, q0 + D, q0 + 2 * D};
                const long long rs[3] = {3LL * D, 3LL * D, 3LL * D};
                const float sc[3] = {p.scale, 1.f, 1.f};
                const int nr[3] = {n, lne, lne};
                wt_load_tiles<3>(tb, toff, src, rs, sc, sm.sTok + buf * 128, nr, tl);
            }
            fence_proxy_async();
            mbar_arrive(&sm.full[buf]);
        }
    } else {
        // ------------------------------------------------------------------ compute warpgroups (WT_NS threads per row)
        const int r = tid & 127, h = tid >> 7;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int nkq = n >> 4;                       // k-steps over the QUERIES (dV = P^T dO, dK = dS^T Q)
        const int li = r >> 4, i1 = (r >> 2) & 3, i2 = r & 3;
        const int bbase = ((li + p.L - 1) * WT_S2 + (i1 + WT_W - 1)) * WT_S2 + (i2 + WT_W - 1) - h * WT_S2 * WT_S2;
        const uint32_t tacc = trow + 256;   // backward: sum_windows dS (bias gradient); this thread's groups at + 16 (WT_NS k + h)
        float* exM = sm.sEx;                       // [NS][128] row max
        float* exS = sm.sEx + WT_NS * 128;         // row sum
        float* exD = sm.sEx + 2 * WT_NS * 128;     // D
        int ne = n;                                   // effective key count of the current scene (wt_keys)
        if (BWD) {
            float z[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0.f;
#pragma unroll
            for (int k = 0; k < WT_GPT; ++k) tmem_st_32x16(tacc + 16 * (WT_NS * k + h), z);
            tmem_st_wait();
        }
        uint32_t ph = 0;
        int it = 0;
        int b_seen = -1;
        uint32_t kmask = 0;
        for (int win = blockIdx.x / p.heads; win < num_windows; win += G, ++it) {
            const int buf = it & 1;
            uint8_t* tb = sm.base + buf * BUF_TILES * TS;
            const uint32_t tb_a = smem_u32(tb);
            const int b = win / (Y * X);
            if (b != b_seen) {          // L dependent global loads: once per scene, not per window (forward 0.89 -> 0.82 ms)
                kmask = wt_kmask(p, b);
                ne = wt_keys(kmask, p.L);
                b_seen = b;
            }
            const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)ne, 0, 0);
            const int nk = ne >> 4;                   // k-steps over the KEYS (O = P V, dQ = dS K)
            mbar_wait(&sm.full[buf], (it >> 1) & 1);
            const long long tok = r < n ? sm.sTok[buf * 128 + r] : 0;
            if (tid == 0) {
                tc_fence_after();
                wt_mma_qk(tmem, desc_lo_word(tb_a, 16), desc_lo_word(tb_a + TS, 16), idesc_s);                         // S
                if (BWD) wt_mma_qk(tmem + 128, desc_lo_word(tb_a + 3 * TS, 16), desc_lo_word(tb_a + 2 * TS, 16), idesc_s);   // dP = dO V^T
                umma_commit(sm.mma);
            }
            mbar_wait(sm.mma, ph);
            ph ^= 1;
            tc_fence_after();
            float s[WT_CPT];
            uint32_t gv;      // bit k: key group WT_NS k + h holds valid keys (the others: masked agent, or beyond ne)
            exM[h * 128 + r] = wt_scores(trow, ne, r, h, kmask, sm.sB, p.L, s, gv);
            wt_bar();
            float m = exM[r];
#pragma unroll
            for (int k = 1; k < WT_NS; ++k) m = fmaxf(m, exM[k * 128 + r]);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < WT_GPT; ++k) {
                if ((gv >> k) & 1u) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        s[16 * k + i] = __expf(s[16 * k + i] - m);
                        sum += s[16 * k + i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) s[16 * k + i] = 0.f;
                }
            }
            exS[h * 128 + r] = sum;
            if (!BWD) {
                // every thread has read its S columns and the S MMAs have retired: P may overwrite Q / K, O may overwrite S
                if (r < n) {   // tiles have n rows: no row >= n; masked groups below ne store their zeros (the MMA reads them)
#pragma unroll
                    for (int k = 0; k < WT_GPT; ++k) {
                        const int g = WT_NS * k + h;
                        if (16 * g < ne) {
                            wt_store8_planes(tb, TS, r, 2 * g, s + 16 * k);
                            wt_store8_planes(tb, TS, r, 2 * g + 1, s + 16 * k + 8);
                        }
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                wt_bar();
                if (tid == 0) {
                    tc_fence_after();
                    wt_mma_pv(tmem, tb_a, tb_a + 4 * TS, TS, nk, 0);
                    umma_commit(sm.mma);
                    umma_commit(&sm.empty[buf]);
                }
                float tot = exS[r];
#pragma unroll
                for (int k = 1; k < WT_NS; ++k) tot += exS[k * 128 + r];
                const float inv = 1.f / tot;
                mbar_wait(sm.mma, ph);
                ph ^= 1;
                tc_fence_after();
                float oh[WT_DPT], ol[WT_DPT];  // features DPT h .. DPT h + DPT - 1: (P Vh) and (P Vl) halves of the accumulator
                wt_tmem_ld_dpt(trow + WT_DPT * h, oh);
                wt_tmem_ld_dpt(trow + 32 + WT_DPT * h, ol);
                tmem_ld_wait();
                if (r < n) {
                    const long long off = tok * D + head * WT_DH + WT_DPT * h;
#pragma unroll
                    for (int c = 0; c < WT_DPT; c += 4)
                        store_split4(p.out, off + c, make_float4((oh[c] + ol[c]) * inv, (oh[c + 1] + ol[c + 1]) * inv,
                                                                 (oh[c + 2] + ol[c + 2]) * inv, (oh[c + 3] + ol[c + 3]) * inv));
                }
            } else {
                wt_bar();
                float tot = exS[r];
#pragma unroll
                for (int k = 1; k < WT_NS; ++k) tot += exS[k * 128 + r];
                const float inv = 1.f / tot;
#pragma unroll
                for (int j = 0; j < WT_CPT; ++j) s[j] *= inv;                    // P
                if (r < n) {
#pragma unroll
                    for (int k = 0; k < WT_GPT; ++k) {
                        const int g = WT_NS * k + h;
                        if (16 * g < ne) {      // masked groups below ne store their zeros (the MMAs read them)
                            wt_store8_planes(tP, TS, r, 2 * g, s + 16 * k);
                            wt_store8_planes(tP, TS, r, 2 * g + 1, s + 16 * k + 8);
                        }
                    }
                }
                float Dv = 0.f;                                                   // this share of D_i = sum_j P_ij dP_ij
#pragma unroll
                for (int k = 0; k < WT_GPT; ++k) {
                    if ((gv >> k) & 1u) {       // P = 0 elsewhere; dP columns of masked / absent groups are never read
                        float dp[16];
                        tmem_ld_32x16(trow + 128 + 16 * (WT_NS * k + h), dp);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) Dv = fmaf(s[16 * k + i], dp[i], Dv);
                    }
                }
                exD[h * 128 + r] = Dv;
                fence_proxy_async();
                tc_fence_before();
                wt_bar();
                if (tid == 0) {
                    tc_fence_after();
                    wt_mma_pv(tmem, smem_u32(tP), tb_a + 3 * TS, TS, nkq, 1);     // dV = P^T dO  -> columns [0, 64) (S is in registers)
                    umma_commit(sm.mma);
                }
                Dv = exD[r];
#pragma unroll
                for (int k = 1; k < WT_NS; ++k) Dv += exD[k * 128 + r];
                mbar_wait(sm.mma, ph);                                            // dV done: the planes may take dS
                ph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < WT_GPT; ++k) {
                    const int g = WT_NS * k + h;
                    if ((gv >> k) & 1u) {
                        float dp[16], ac[16];
                        tmem_ld_32x16(trow + 128 + 16 * g, dp);
                        tmem_ld_32x16(tacc + 16 * g, ac);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float ds = s[16 * k + i] * (dp[i] - Dv);         // dS_ij
                            dp[i] = ds;
                            ac[i] += ds;
                        }
                        tmem_st_32x16(tacc + 16 * g, ac);
                        if (r < n) {
                            wt_store8_planes(tP, TS, r, 2 * g, dp);
                            wt_store8_planes(tP, TS, r, 2 * g + 1, dp + 8);
                        }
                    } else if (16 * g < ne && r < n) {                          // a masked agent among the valid ones: dS = 0
                        float zz[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) zz[i] = 0.f;
                        wt_store8_planes(tP, TS, r, 2 * g, zz);
                        wt_store8_planes(tP, TS, r, 2 * g + 1, zz);
                    }
                }
                tmem_st_wait();
                fence_proxy_async();
                tc_fence_before();
                wt_bar();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t pa = smem_u32(tP);
                    wt_mma_pv(tmem + 64, pa, tb_a, TS, nkq, 1);                   // dK = dS^T (scale Q)
                    wt_mma_pv(tmem + 128, pa, tb_a + TS, TS, nk, 0);              // dQ = dS K (scaled below); dP has been consumed
                    umma_commit(sm.mma);
                    umma_commit(&sm.empty[buf]);
                }
                mbar_wait(sm.mma, ph);
                ph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int part = 0; part < 3; ++part) {                            // 0: dV, 1: dK, 2: dQ
                    float oh[WT_DPT], ol[WT_DPT];
                    wt_tmem_ld_dpt(trow + part * 64 + WT_DPT * h, oh);
                    wt_tmem_ld_dpt(trow + part * 64 + 32 + WT_DPT * h, ol);
                    tmem_ld_wait();
                    if (r < n) {
                        const float f = part == 2 ? p.scale : 1.f;
                        const bool keep = part == 2 || r < ne;   // dV / dK rows of the trailing masked keys: zero (their accumulator rows are stale)
                        const long long off = tok * (3 * D) + (2 - part) * D + head * WT_DH + WT_DPT * h;
#pragma unroll
                        for (int c = 0; c < WT_DPT; c += 4)
                            store_split4(p.out, off + c,
                                         keep ? make_float4((oh[c] + ol[c]) * f, (oh[c + 1] + ol[c + 1]) * f,
                                                            (oh[c + 2] + ol[c + 2]) * f, (oh[c + 3] + ol[c + 3]) * f)
                                              : make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
            }
            tc_fence_before();
            wt_bar();   // the accumulators, the exchange slots (and, backward, the planes) are reused by the next window
        }
        if (BWD) {
            // fold the per-thread columns of sum dS into the head's table (entry of (i, j) = bbase_i - sub_j)
#pragma unroll
            for (int k = 0; k < WT_GPT; ++k) {
                const int g = WT_NS * k + h;
                if (16 * g < n) {
                    float ac[16];
                    tmem_ld_32x16(tacc + 16 * g, ac);
                    tmem_ld_wait();
                    if (r < n) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (ac[i] != 0.f)
                                atomicAdd(&sm.sdB[bbase - (((WT_NS * k) * WT_S2 + ((i >> 2) & 3)) * WT_S2 + (i & 3))], ac[i]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (BWD)
        for (int i = tid; i < nb; i += WT_THREADS)
            if (sm.sdB[i] != 0.f) atomicAdd(&p.dbias[i * p.heads + head], sm.sdB[i]);
    if (warp == 0) tmem_dealloc<TM_COLS>(tmem);
}

}  // namespace a2x
