// V2X-ViT fusion kernels that are not plain linears / LayerNorm / window attention (those are shared with CoBEVT):
// relative temporal encoding, heterogeneous multi-agent attention (HGT) with the relation tensors folded into the K / V
// projections, and the split-attention fusion of the pyramid window branches.
// Token tensors are fp32 NHWC [agents][H][W][C] of ONE padded-free scene batch (padded agents are never keys and only
// agent 0 is returned, so they are skipped: exact).
//
// Reference semantics:
//   opencood/models/v2xvit_modules/v2xvit_basic.py:41-80    RelTemporalEncoding / RTE
//   opencood/models/v2xvit_modules/hmsa.py:37-158           HGTCavAttention (typed q/k/v/a linears, relation_att/msg)
//   opencood/models/v2xvit_modules/split_attn.py:6-63       RadixSoftmax / SplitAttn
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"

namespace a2x {

// ---------------------------------------------------------------------------------------------- RTE
// vec[a][c] = sum_k W[c][k] * emb[idx[a]][k] + b[c]           (one block per agent)
__global__ void rte_vectors_kernel(const float* __restrict__ emb, const int* __restrict__ idx,
                                   const float* __restrict__ W, const float* __restrict__ b, int C,
                                   float* __restrict__ vec) {
    extern __shared__ float se[];
    const int a = blockIdx.x;
    const float* e = emb + (long long)idx[a] * C;
    for (int k = threadIdx.x; k < C; k += blockDim.x) se[k] = e[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < C; ++k) s = fmaf(W[(long long)c * C + k], se[k], s);
        vec[(long long)a * C + c] = s + b[c];
    }
}

// x[a][p][c] += vec[a][c]
__global__ void agent_vec_add_kernel(float* __restrict__ x, const float* __restrict__ vec, long long pix, int C,
                                     long long total4) {
    const int q = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % q) * 4;
        const long long a = i / (q * pix);
        float4 v = reinterpret_cast<float4*>(x)[i];
        const float4 d = *reinterpret_cast<const float4*>(vec + a * C + c);
        v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
        reinterpret_cast<float4*>(x)[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------- HGT weight folding
// For key-agent type tj the fused projection has 5*C output rows:
//   [0,C)   q  = Wq[tj]
//   [C,2C)  k' for query type 0 = A[0*T+tj] Wk[tj]      [2C,3C) k' for query type 1 = A[1*T+tj] Wk[tj]
//   [3C,4C) v' for query type 0 = M[0*T+tj]^T Wv[tj]    [4C,5C) v' for query type 1 = M[1*T+tj]^T Wv[tj]
// with A = relation_att, M = relation_msg ([relations][heads][dh][dh], block diagonal over heads), so that
//   logit(i,j) = q_i . k'_{type_i}(j)   and   message(i,j) = v'_{type_i}(j)          (hmsa.py:136-151 re-associated).
struct HgtFoldParams {
    const float* qw[2]; const float* qb[2];
    const float* kw[2]; const float* kb[2];
    const float* vw[2]; const float* vb[2];
    const float* rel_att; const float* rel_msg;
    float* wf;  // [2][5C][C]
    float* bf;  // [2][5C]
    int C, heads, dh;
};

__global__ void hgt_fold_kernel(const HgtFoldParams p) {
    const int C = p.C, dh = p.dh;
    const long long per_type = (long long)5 * C * (C + 1);  // weights then one bias column
    const long long total = 2 * per_type;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int tj = (int)(i / per_type);
        long long r = i - tj * per_type;
        const int col = (int)(r % (C + 1));  // col == C -> bias
        const int row = (int)(r / (C + 1));
        const int part = row / C, o = row - part * C;
        float val;
        if (part == 0) {
            val = col < C ? p.qw[tj][(long long)o * C + col] : p.qb[tj][o];
        } else {
            const int ti = (part - 1) & 1;
            const bool is_v = part >= 3;
            const int h = o / dh, a = o - h * dh;
            const float* R = (is_v ? p.rel_msg : p.rel_att) + ((long long)(ti * 2 + tj) * p.heads + h) * dh * dh;
            const float* Wm = is_v ? p.vw[tj] : p.kw[tj];
            const float* Bv = is_v ? p.vb[tj] : p.kb[tj];
            float s = 0.f;
            for (int t = 0; t < dh; ++t) {
                const float rv = is_v ? R[t * dh + a] : R[a * dh + t];  // v' = M^T v ; k' = A k
                const float src = col < C ? Wm[(long long)(h * dh + t) * C + col] : Bv[h * dh + t];
                s = fmaf(rv, src, s);
            }
            val = s;
        }
        if (col < C) p.wf[((long long)tj * 5 * C + row) * C + col] = val;
        else p.bf[(long long)tj * 5 * C + row] = val;
    }
}

// ---------------------------------------------------------------------------------------------- HGT attention
// one thread per (pixel, query agent i, head m): softmax_j( scale * q_i . k'_{ti}(j) ) over the keys valid at this
// pixel (mask[j][p] != 0), out_i = sum_j att * v'_{ti}(j).   qkv: [n][pix][5C] (layout above).
template <int DH>
__global__ void __launch_bounds__(256) hgt_attention_kernel(const float* __restrict__ qkv, const int* __restrict__ types,
                                                            const float* __restrict__ mask, int n, long long pix,
                                                            int heads, float scale, SplitOut out) {
    const int C = heads * DH;
    const long long total = pix * n * heads;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(t % heads);
        const int i = (int)((t / heads) % n);
        const long long p = t / ((long long)heads * n);
        const int ti = types[i];
        float q[DH], acc[DH];
        const float* qrow = qkv + ((long long)i * pix + p) * 5 * C + m * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 v = *reinterpret_cast<const float4*>(qrow + c);
            q[c] = v.x * scale; q[c + 1] = v.y * scale; q[c + 2] = v.z * scale; q[c + 3] = v.w * scale;
        }
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = 0.f;
        float mx = -INFINITY, den = 0.f;
        for (int j = 0; j < n; ++j) {
            if (mask[(long long)j * pix + p] == 0.f) continue;
            const float* krow = qkv + ((long long)j * pix + p) * 5 * C + (1 + ti) * C + m * DH;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(krow + c);
                s0 = fmaf(q[c], v.x, s0); s1 = fmaf(q[c + 1], v.y, s1);
                s0 = fmaf(q[c + 2], v.z, s0); s1 = fmaf(q[c + 3], v.w, s1);
            }
            const float s = s0 + s1;
            const float mn = fmaxf(mx, s);
            const float corr = __expf(mx - mn), e = __expf(s - mn);
            den = den * corr + e;
            const float* vrow = krow + 2 * C;
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(vrow + c);
                acc[c] = fmaf(acc[c], corr, e * v.x); acc[c + 1] = fmaf(acc[c + 1], corr, e * v.y);
                acc[c + 2] = fmaf(acc[c + 2], corr, e * v.z); acc[c + 3] = fmaf(acc[c + 3], corr, e * v.w);
            }
            mx = mn;
        }
        const float inv = 1.f / den;  // no valid key -> NaN, like softmax over an all -inf row in the reference
        const long long off = ((long long)i * pix + p) * C + m * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4)
            store_split4(out, off + c, make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv));
    }
}

// Cooperative variant (the one launched): a GROUP of DH / 4 lanes owns one (pixel, head) and every lane 4 of its DH
// features. The thread-per-(pixel, agent, head) kernel above fetches a 128-byte row per thread with float4 loads — 32
// different lines per request: ncu showed the L1 data pipe at 95 % of peak with the issue slots 8 % active and DRAM at a third
// of its rate. Here the group stages the pixel's 5 n rows (q, k'(0), k'(1), v'(0), v'(1) of every agent: each read ONCE,
// consecutive lanes = consecutive 16-byte chunks) in shared memory, dot products are reduced with log2(DH / 4) shuffles,
// and every output row is stored by the group as one contiguous line.
template <int DH>
__global__ void __launch_bounds__(128) hgt_attention_coop_kernel(const float* __restrict__ qkv, const int* __restrict__ types,
                                                                 const float* __restrict__ mask, int n, long long pix,
                                                                 int heads, float scale, SplitOut out) {
    constexpr int LPG = DH / 4;          // lanes per group
    constexpr int GPB = 128 / LPG;       // groups per block
    extern __shared__ float hsm[];       // [GPB][5 n][DH]
    const int C = heads * DH;
    const int g = threadIdx.x / LPG, s = threadIdx.x % LPG;
    const long long G = (long long)blockIdx.x * GPB + g;
    const int m = (int)(G % heads);
    const long long p = G / heads;
    const bool valid = p < pix;
    float* my = hsm + (long long)g * (5 * n * DH);
    const unsigned gmask = (LPG == 32 ? 0xffffffffu : ((1u << LPG) - 1u)) << ((threadIdx.x & 31) / LPG * LPG);
    if (valid) {
        for (int r = 0; r < 5 * n; ++r) {
            const int a = r / 5, slot = r - a * 5;
            *reinterpret_cast<float4*>(my + r * DH + 4 * s) =
                *reinterpret_cast<const float4*>(qkv + ((long long)a * pix + p) * 5 * C + slot * C + m * DH + 4 * s);
        }
    }
    __syncwarp(gmask);
    if (!valid) return;
    for (int i = 0; i < n; ++i) {
        const int ti = types[i];
        float4 q = *reinterpret_cast<const float4*>(my + (i * 5) * DH + 4 * s);
        q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float mx = -INFINITY, den = 0.f;
        for (int j = 0; j < n; ++j) {
            if (mask[(long long)j * pix + p] == 0.f) continue;       // group-uniform
            const float4 k = *reinterpret_cast<const float4*>(my + (j * 5 + 1 + ti) * DH + 4 * s);
            float d = fmaf(q.x, k.x, fmaf(q.y, k.y, fmaf(q.z, k.z, q.w * k.w)));
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) d += __shfl_xor_sync(gmask, d, o);
            const float mn = fmaxf(mx, d);
            const float corr = __expf(mx - mn), e = __expf(d - mn);
            den = den * corr + e;
            const float4 v = *reinterpret_cast<const float4*>(my + (j * 5 + 3 + ti) * DH + 4 * s);
            acc.x = fmaf(acc.x, corr, e * v.x); acc.y = fmaf(acc.y, corr, e * v.y);
            acc.z = fmaf(acc.z, corr, e * v.z); acc.w = fmaf(acc.w, corr, e * v.w);
            mx = mn;
        }
        const float inv = 1.f / den;  // no valid key -> NaN, like softmax over an all -inf row in the reference
        store_split4(out, ((long long)i * pix + p) * C + m * DH + 4 * s, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
    }
}

// ---------------------------------------------------------------------------------------------- split attention
// sums[a][c] += sum_p (w0 + w1 + w2)[a][p][c]      (global average pool numerator)
__global__ void __launch_bounds__(256) split_pool_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                                         const float* __restrict__ w2, long long pix, int C,
                                                         int chunks, float* __restrict__ partials) {
    // grid = (chunks, agents); blockDim = C/4 * rows. Every block stores its partial sum [C] at partials[a][chunk]: no
    // atomics, so that the pooled vector (reduced in chunk order by split_weights_kernel) is bit-reproducible
    const int q = C >> 2;
    const int cq = threadIdx.x % q, prow = threadIdx.x / q, ppb = blockDim.x / q;
    const int a = blockIdx.y;
    const long long per = (pix + chunks - 1) / chunks;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(pix, p0 + per);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prow < ppb) {
        for (long long p = p0 + prow; p < p1; p += ppb) {
            const long long o = ((long long)a * pix + p) * C + cq * 4;
            const float4 x = *reinterpret_cast<const float4*>(w0 + o);
            const float4 y = *reinterpret_cast<const float4*>(w1 + o);
            const float4 z = *reinterpret_cast<const float4*>(w2 + o);
            s.x += (x.x + y.x) + z.x; s.y += (x.y + y.y) + z.y; s.z += (x.z + y.z) + z.z; s.w += (x.w + y.w) + z.w;
        }
    }
    extern __shared__ float4 red4[];
    red4[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < q) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < ppb; ++r) {
            const float4 v = red4[r * q + threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        *reinterpret_cast<float4*>(partials + ((long long)a * chunks + blockIdx.x) * C + threadIdx.x * 4) = t;
    }
}

// per agent: g = sums / pix; h = relu(LN(fc1 g)); a = fc2 h ([3C]); wts[r][c] = softmax_r a[r*C + c]
__global__ void split_weights_kernel(const float* __restrict__ partials, int chunks, float* __restrict__ sums, float inv_pix,
                                     const float* __restrict__ fc1, const float* __restrict__ ln_g,
                                     const float* __restrict__ ln_b, const float* __restrict__ fc2, int C,
                                     float* __restrict__ wts) {
    extern __shared__ float sm[];
    float* g = sm;          // [C]
    float* h = sm + C;      // [C]
    float* red = h + C;     // [2]
    const int a = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;      // fixed chunk order: deterministic
        for (int k = 0; k < chunks; ++k) s += partials[((long long)a * chunks + k) * C + c];
        sums[(long long)a * C + c] = s;   // kept for the backward
        g[c] = s * inv_pix;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < C; ++k) s = fmaf(fc1[(long long)c * C + k], g[k], s);
        h[c] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // C <= 512: a serial two-pass LayerNorm is negligible
        float m = 0.f;
        for (int c = 0; c < C; ++c) m += h[c];
        m /= C;
        float v = 0.f;
        for (int c = 0; c < C; ++c) v += (h[c] - m) * (h[c] - m);
        red[0] = m;
        red[1] = rsqrtf(v / C + 1e-5f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = (h[c] - red[0]) * red[1] * ln_g[c] + ln_b[c];
        g[c] = fmaxf(v, 0.f);  // reuse g for the activated hidden vector (all reads of g finished above)
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float t = 0.f;
            const float* w = fc2 + (long long)(r * C + c) * C;
            for (int k = 0; k < C; ++k) t = fmaf(w[k], g[k], t);
            s[r] = t;
        }
        const float mx = fmaxf(s[0], fmaxf(s[1], s[2]));
        const float e0 = expf(s[0] - mx), e1 = expf(s[1] - mx), e2 = expf(s[2] - mx);
        const float inv = 1.f / (e0 + e1 + e2);
        float* o = wts + (long long)a * 3 * C;
        o[c] = e0 * inv;
        o[C + c] = e1 * inv;
        o[2 * C + c] = e2 * inv;
    }
}

// x[a][p][c] += w0*wts[a][0][c] + w1*wts[a][1][c] + w2*wts[a][2][c]      (PWA output + residual)
__global__ void __launch_bounds__(256) split_combine_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                                            const float* __restrict__ w2, const float* __restrict__ wts,
                                                            float* __restrict__ x, long long pix, int C,
                                                            long long total4) {
    const int q = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % q) * 4;
        const long long a = i / (q * pix);
        const float* wa = wts + a * 3 * C + c;
        const float4 a0 = *reinterpret_cast<const float4*>(wa);
        const float4 a1 = *reinterpret_cast<const float4*>(wa + C);
        const float4 a2 = *reinterpret_cast<const float4*>(wa + 2 * C);
        const float4 u = reinterpret_cast<const float4*>(w0)[i];
        const float4 v = reinterpret_cast<const float4*>(w1)[i];
        const float4 w = reinterpret_cast<const float4*>(w2)[i];
        float4 r = reinterpret_cast<float4*>(x)[i];
        r.x += u.x * a0.x + v.x * a1.x + w.x * a2.x;
        r.y += u.y * a0.y + v.y * a1.y + w.y * a2.y;
        r.z += u.z * a0.z + v.z * a1.z + w.z * a2.z;
        r.w += u.w * a0.w + v.w * a1.w + w.w * a2.w;
        reinterpret_cast<float4*>(x)[i] = r;
    }
}

static int vx_grid(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_rte_add(float* x, int n_agents, long long pix, int C, const float* emb_table, const int* emb_idx_dev,
                const float* lin_w, const float* lin_b, float* vec_ws, a2x_stream_t stream) {
    A2X_REQUIRE(x && emb_table && emb_idx_dev && lin_w && lin_b && vec_ws && n_agents > 0 && pix > 0 && C > 0 && C % 4 == 0,
                "rte_add: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    rte_vectors_kernel<<<n_agents, 256, C * sizeof(float), st>>>(emb_table, emb_idx_dev, lin_w, lin_b, C, vec_ws);
    A2X_LAUNCHED();
    const long long total4 = (long long)n_agents * pix * (C / 4);
    agent_vec_add_kernel<<<vx_grid(total4), 256, 0, st>>>(x, vec_ws, pix, C, total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_hgt_fold(const float* const* qw, const float* const* qb, const float* const* kw, const float* const* kb,
                 const float* const* vw, const float* const* vb, const float* relation_att, const float* relation_msg,
                 int C, int heads, float* w_fold, float* b_fold, a2x_stream_t stream) {
    A2X_REQUIRE(qw && qb && kw && kb && vw && vb && relation_att && relation_msg && w_fold && b_fold && C > 0 &&
                    heads > 0 && C % heads == 0,
                "hgt_fold: bad args (two agent types expected)");
    HgtFoldParams p;
    for (int t = 0; t < 2; ++t) {
        p.qw[t] = qw[t]; p.qb[t] = qb[t]; p.kw[t] = kw[t]; p.kb[t] = kb[t]; p.vw[t] = vw[t]; p.vb[t] = vb[t];
    }
    p.rel_att = relation_att; p.rel_msg = relation_msg; p.wf = w_fold; p.bf = b_fold;
    p.C = C; p.heads = heads; p.dh = C / heads;
    hgt_fold_kernel<<<vx_grid((long long)2 * 5 * C * (C + 1)), 256, 0, (cudaStream_t)stream>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_hgt_attention_fwd(const float* qkv, const int* types_dev, const float* key_mask, int n_agents, long long pix,
                          int heads, int dim_head, float scale, const a2x_output* out, a2x_stream_t stream) {
    A2X_REQUIRE(qkv && types_dev && key_mask && out && out->hi && n_agents > 0 && pix > 0 && heads > 0,
                "hgt_attention_fwd: bad args");
    A2X_REQUIRE(out->cs == heads * dim_head, "hgt_attention_fwd: dense [.., heads*dim_head] output expected");
    SplitOut o;
    o.hi = out->hi; o.b16 = (__nv_bfloat16*)out->b16; o.ps = out->b16_plane;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t csm = (size_t)(128 / (dim_head / 4)) * 5 * n_agents * dim_head * sizeof(float);   // [groups][5 n][DH]
    if (csm <= 200 * 1024 && (dim_head == 16 || dim_head == 32 || dim_head == 64)) {
        const long long groups = pix * heads, gpb = 128 / (dim_head / 4);
        const unsigned blocks = (unsigned)((groups + gpb - 1) / gpb);
#define A2X_HGT(DH)                                                                                                   \
    do {                                                                                                              \
        A2X_CHECK_CUDA(cudaFuncSetAttribute(hgt_attention_coop_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm)); \
        hgt_attention_coop_kernel<DH><<<blocks, 128, csm, st>>>(qkv, types_dev, key_mask, n_agents, pix, heads, scale, o); \
    } while (0)
        if (dim_head == 16) A2X_HGT(16);
        else if (dim_head == 32) A2X_HGT(32);
        else A2X_HGT(64);
#undef A2X_HGT
        A2X_LAUNCHED();
        A2X_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const int g = vx_grid(pix * n_agents * heads);
    if (dim_head == 16) hgt_attention_kernel<16><<<g, 256, 0, st>>>(qkv, types_dev, key_mask, n_agents, pix, heads, scale, o);
    else if (dim_head == 32) hgt_attention_kernel<32><<<g, 256, 0, st>>>(qkv, types_dev, key_mask, n_agents, pix, heads, scale, o);
    else if (dim_head == 64) hgt_attention_kernel<64><<<g, 256, 0, st>>>(qkv, types_dev, key_mask, n_agents, pix, heads, scale, o);
    else {
        set_error("hgt_attention_fwd: dim_head %d not in {16, 32, 64}", dim_head);
        return 1;
    }
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_split_attn_fuse(const float* w0, const float* w1, const float* w2, int n_agents, long long pix, int C,
                        const float* fc1, const float* ln_gamma, const float* ln_beta, const float* fc2,
                        float* sums_ws, float* partials_ws, float* weights_ws, float* x_inout, a2x_stream_t stream) {
    A2X_REQUIRE(w0 && w1 && w2 && fc1 && ln_gamma && ln_beta && fc2 && sums_ws && partials_ws && weights_ws && x_inout &&
                    n_agents > 0 && pix > 0 && C > 0 && C % 4 == 0 && C <= 1024 && 256 % (C / 4) == 0,
                "split_attn_fuse: bad args (C/4 must divide 256)");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = A2X_SPLIT_ATTN_CHUNKS;   // fixed: the pooled sums do not depend on the number of agents in the call
    split_pool_kernel<<<dim3(chunks, n_agents), 256, 256 * sizeof(float4), st>>>(w0, w1, w2, pix, C, chunks, partials_ws);
    A2X_LAUNCHED();
    split_weights_kernel<<<n_agents, 256, (2 * C + 2) * sizeof(float), st>>>(partials_ws, chunks, sums_ws, 1.0f / (float)pix,
                                                                           fc1, ln_gamma, ln_beta, fc2, C, weights_ws);
    A2X_LAUNCHED();
    const long long total4 = (long long)n_agents * pix * (C / 4);
    split_combine_kernel<<<vx_grid(total4), 256, 0, st>>>(w0, w1, w2, weights_ws, x_inout, pix, C, total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

// =====================================================================================================================
// Backward kernels of the V2X-ViT fusion (training step): autograd of HGTCavAttention (through the folded projections),
// SplitAttn and RTE. Same re-association as the forward; everything is recomputed from the saved projections.
namespace a2x {

constexpr int HGT_MAX_AGENTS = 12;  // backward: 2 n^2 floats of shared memory per thread

// one thread per (pixel, head); dqkv has the layout of qkv: q | k'(0) | k'(1) | v'(0) | v'(1), every slot written here.
// Phase A, per query agent i: probability row P_i. and dS_i. -> the thread's shared-memory slab, dq_i from registers.
// Phase B, per key agent j and query type ty: dk'_ty(j) = sum_{i of type ty} dS_ij q_i, dv'_ty(j) = sum P_ij dO_i,
// accumulated in registers and stored once (no read-modify-write of global memory).
template <int DH>
__global__ void __launch_bounds__(128) hgt_attention_bwd_kernel(const float* __restrict__ qkv, const int* __restrict__ types,
                                                                const float* __restrict__ mask,
                                                                const float* __restrict__ dout, int n, long long pix,
                                                                int heads, float scale, float* __restrict__ dqkv) {
    extern __shared__ float hsm[];           // [2][n][n][128]: P, dS (thread-interleaved: conflict free)
    const int C = heads * DH;
    const long long total = pix * heads;
    float* sP = hsm + threadIdx.x;
    float* sS = sP + n * n * 128;
    auto ld = [&](const float* src, float (&v)[DH], float mul) {
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(src + c);
            v[c] = t.x * mul; v[c + 1] = t.y * mul; v[c + 2] = t.z * mul; v[c + 3] = t.w * mul;
        }
    };
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(t % heads);
        const long long p = t / heads;
        for (int i = 0; i < n; ++i) {
            const int ti = types[i];
            float q[DH], go[DH];
            ld(qkv + ((long long)i * pix + p) * 5 * C + m * DH, q, scale);
            ld(dout + ((long long)i * pix + p) * C + m * DH, go, 1.f);
            float mx = -INFINITY;
            for (int j = 0; j < n; ++j) {
                float a = -INFINITY, bsum = 0.f;
                if (mask[(long long)j * pix + p] != 0.f) {
                    const float* krow = qkv + ((long long)j * pix + p) * 5 * C + (1 + ti) * C + m * DH;
                    const float* vrow = krow + 2 * C;
                    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 k4 = *reinterpret_cast<const float4*>(krow + c);
                        const float4 v4 = *reinterpret_cast<const float4*>(vrow + c);
                        a0 = fmaf(q[c], k4.x, a0); a1 = fmaf(q[c + 1], k4.y, a1);
                        a0 = fmaf(q[c + 2], k4.z, a0); a1 = fmaf(q[c + 3], k4.w, a1);
                        b0 = fmaf(go[c], v4.x, b0); b1 = fmaf(go[c + 1], v4.y, b1);
                        b0 = fmaf(go[c + 2], v4.z, b0); b1 = fmaf(go[c + 3], v4.w, b1);
                    }
                    a = a0 + a1;
                    bsum = b0 + b1;
                    mx = fmaxf(mx, a);
                }
                sP[(i * n + j) * 128] = a;
                sS[(i * n + j) * 128] = bsum;
            }
            float l = 0.f;
            for (int j = 0; j < n; ++j) {
                const float e = expf(sP[(i * n + j) * 128] - mx);
                sP[(i * n + j) * 128] = e;
                l += e;
            }
            const float inv = 1.f / l;
            float Dv = 0.f;
            for (int j = 0; j < n; ++j) {
                const float pj = sP[(i * n + j) * 128] * inv;
                sP[(i * n + j) * 128] = pj;
                Dv = fmaf(pj, sS[(i * n + j) * 128], Dv);
            }
            float dq[DH];
#pragma unroll
            for (int c = 0; c < DH; ++c) dq[c] = 0.f;
            for (int j = 0; j < n; ++j) {
                const float pj = sP[(i * n + j) * 128];
                const float ds = pj * (sS[(i * n + j) * 128] - Dv);
                sS[(i * n + j) * 128] = ds;
                if (pj == 0.f) continue;
                const float* krow = qkv + ((long long)j * pix + p) * 5 * C + (1 + ti) * C + m * DH;
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(krow + c);
                    dq[c] = fmaf(ds, k4.x, dq[c]); dq[c + 1] = fmaf(ds, k4.y, dq[c + 1]);
                    dq[c + 2] = fmaf(ds, k4.z, dq[c + 2]); dq[c + 3] = fmaf(ds, k4.w, dq[c + 3]);
                }
            }
            float* dqo = dqkv + ((long long)i * pix + p) * 5 * C + m * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 4)
                *reinterpret_cast<float4*>(dqo + c) = make_float4(dq[c] * scale, dq[c + 1] * scale, dq[c + 2] * scale, dq[c + 3] * scale);
        }
        for (int j = 0; j < n; ++j) {
            for (int ty = 0; ty < 2; ++ty) {
                float dk[DH], dv[DH];
#pragma unroll
                for (int c = 0; c < DH; ++c) dk[c] = dv[c] = 0.f;
                for (int i = 0; i < n; ++i) {
                    if (types[i] != ty) continue;
                    const float pj = sP[(i * n + j) * 128];
                    if (pj == 0.f) continue;
                    const float ds = sS[(i * n + j) * 128] * scale;   // s = (q * scale) . k'
                    const float* qrow = qkv + ((long long)i * pix + p) * 5 * C + m * DH;
                    const float* grow = dout + ((long long)i * pix + p) * C + m * DH;
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 q4 = *reinterpret_cast<const float4*>(qrow + c);
                        const float4 g4 = *reinterpret_cast<const float4*>(grow + c);
                        dk[c] = fmaf(ds, q4.x, dk[c]); dk[c + 1] = fmaf(ds, q4.y, dk[c + 1]);
                        dk[c + 2] = fmaf(ds, q4.z, dk[c + 2]); dk[c + 3] = fmaf(ds, q4.w, dk[c + 3]);
                        dv[c] = fmaf(pj, g4.x, dv[c]); dv[c + 1] = fmaf(pj, g4.y, dv[c + 1]);
                        dv[c + 2] = fmaf(pj, g4.z, dv[c + 2]); dv[c + 3] = fmaf(pj, g4.w, dv[c + 3]);
                    }
                }
                float* dko = dqkv + ((long long)j * pix + p) * 5 * C + (1 + ty) * C + m * DH;
                float* dvo = dko + 2 * C;
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    *reinterpret_cast<float4*>(dko + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
                    *reinterpret_cast<float4*>(dvo + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
                }
            }
        }
    }
}

// Cooperative variant (the one launched), same layout of work as hgt_attention_coop_kernel: a group of DH / 4 lanes owns
// one (pixel, head), stages the pixel's 6 n rows (q, k'(0), k'(1), v'(0), v'(1), dO of every agent: each read ONCE, coalesced)
// in shared memory, keeps P and dS of the n x n agent pairs in a small per-group slab, reduces the dot products with
// shuffles and stores every gradient row as one contiguous line. Phase A, per query agent i: P_i., dS_i., dq_i. Phase B, per
// key agent j and query type ty: dk'_ty(j) = sum_{i of type ty} dS_ij (scale q_i), dv'_ty(j) = sum P_ij dO_i.
template <int DH>
__global__ void __launch_bounds__(128) hgt_attention_bwd_coop_kernel(const float* __restrict__ qkv, const int* __restrict__ types,
                                                                     const float* __restrict__ mask,
                                                                     const float* __restrict__ dout, int n, long long pix,
                                                                     int heads, float scale, float* __restrict__ dqkv) {
    constexpr int LPG = DH / 4;
    constexpr int GPB = 128 / LPG;
    extern __shared__ float hsm[];       // [GPB][6 n][DH] rows, then [GPB][2][n][n] P | dS
    const int C = heads * DH;
    const int g = threadIdx.x / LPG, s = threadIdx.x % LPG;
    const long long G = (long long)blockIdx.x * GPB + g;
    const int m = (int)(G % heads);
    const long long p = G / heads;
    const bool valid = p < pix;
    float* my = hsm + (long long)g * (6 * n * DH);
    float* sP = hsm + (long long)GPB * (6 * n * DH) + (long long)g * (2 * n * n);
    float* sS = sP + n * n;
    const unsigned gmask = (LPG == 32 ? 0xffffffffu : ((1u << LPG) - 1u)) << ((threadIdx.x & 31) / LPG * LPG);
    if (valid) {
        for (int r = 0; r < 6 * n; ++r) {
            const int a = r / 6, slot = r - a * 6;
            const float* src = slot < 5 ? qkv + ((long long)a * pix + p) * 5 * C + slot * C + m * DH + 4 * s
                                        : dout + ((long long)a * pix + p) * C + m * DH + 4 * s;
            *reinterpret_cast<float4*>(my + r * DH + 4 * s) = *reinterpret_cast<const float4*>(src);
        }
    }
    __syncwarp(gmask);
    if (!valid) return;
    auto row = [&](int a, int slot) { return *reinterpret_cast<const float4*>(my + (a * 6 + slot) * DH + 4 * s); };
    auto gsum = [&](float v) {
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    };
    // ---- phase A
    for (int i = 0; i < n; ++i) {
        const int ti = types[i];
        float4 q = row(i, 0);
        q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;
        const float4 go = row(i, 5);
        float mx = -INFINITY;
        for (int j = 0; j < n; ++j) {
            float a = -INFINITY, bsum = 0.f;
            if (mask[(long long)j * pix + p] != 0.f) {       // group-uniform
                const float4 k = row(j, 1 + ti), v = row(j, 3 + ti);
                a = gsum(fmaf(q.x, k.x, fmaf(q.y, k.y, fmaf(q.z, k.z, q.w * k.w))));
                bsum = gsum(fmaf(go.x, v.x, fmaf(go.y, v.y, fmaf(go.z, v.z, go.w * v.w))));
                mx = fmaxf(mx, a);
            }
            if (s == 0) {
                sP[i * n + j] = a;
                sS[i * n + j] = bsum;
            }
        }
        __syncwarp(gmask);
        float l = 0.f;
        for (int j = 0; j < n; ++j) l += expf(sP[i * n + j] - mx);
        const float inv = 1.f / l;
        float Dv = 0.f;
        for (int j = 0; j < n; ++j) Dv = fmaf(expf(sP[i * n + j] - mx) * inv, sS[i * n + j], Dv);
        float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp(gmask);       // every lane has read the raw scores before lane 0 overwrites them
        for (int j = 0; j < n; ++j) {
            const float pj = expf(sP[i * n + j] - mx) * inv;
            const float ds = pj * (sS[i * n + j] - Dv);
            __syncwarp(gmask);
            if (s == 0) {
                sP[i * n + j] = pj;
                sS[i * n + j] = ds;
            }
            if (pj == 0.f) continue;                          // group-uniform (same values in every lane)
            const float4 k = row(j, 1 + ti);
            dq.x = fmaf(ds, k.x, dq.x); dq.y = fmaf(ds, k.y, dq.y); dq.z = fmaf(ds, k.z, dq.z); dq.w = fmaf(ds, k.w, dq.w);
        }
        *reinterpret_cast<float4*>(dqkv + ((long long)i * pix + p) * 5 * C + m * DH + 4 * s) =
            make_float4(dq.x * scale, dq.y * scale, dq.z * scale, dq.w * scale);
    }
    __syncwarp(gmask);
    // ---- phase B
    for (int j = 0; j < n; ++j) {
        for (int ty = 0; ty < 2; ++ty) {
            float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
            for (int i = 0; i < n; ++i) {
                if (types[i] != ty) continue;
                const float pj = sP[i * n + j];
                if (pj == 0.f) continue;
                const float ds = sS[i * n + j] * scale;       // s = (q * scale) . k'
                const float4 q = row(i, 0), go = row(i, 5);
                dk.x = fmaf(ds, q.x, dk.x); dk.y = fmaf(ds, q.y, dk.y); dk.z = fmaf(ds, q.z, dk.z); dk.w = fmaf(ds, q.w, dk.w);
                dv.x = fmaf(pj, go.x, dv.x); dv.y = fmaf(pj, go.y, dv.y); dv.z = fmaf(pj, go.z, dv.z); dv.w = fmaf(pj, go.w, dv.w);
            }
            float* dko = dqkv + ((long long)j * pix + p) * 5 * C + (1 + ty) * C + m * DH + 4 * s;
            *reinterpret_cast<float4*>(dko) = dk;
            *reinterpret_cast<float4*>(dko + 2 * C) = dv;
        }
    }
}

// Backward of hgt_fold: gradients of the fused projection (dwf [2][5C][C], dbf [2][5C]) -> typed q/k/v linears and the
// relation tensors. Weights and biases are handled as one [C][C+1] extended matrix (last column = bias).
struct HgtFoldBwdParams {
    const float* dwf; const float* dbf;
    const float* kw[2]; const float* kb[2]; const float* vw[2]; const float* vb[2];
    const float* rel_att; const float* rel_msg;
    float* dqw[2]; float* dqb[2]; float* dkw[2]; float* dkb[2]; float* dvw[2]; float* dvb[2];
    float* drel_att; float* drel_msg;
    int C, heads, dh;
};

__global__ void hgt_fold_bwd_kernel(const HgtFoldBwdParams p) {
    const int C = p.C, dh = p.dh, E = C + 1;
    auto dWf = [&](int tj, int row, int col) { return col < C ? p.dwf[((long long)tj * 5 * C + row) * C + col] : p.dbf[tj * 5 * C + row]; };
    const long long n1 = (long long)2 * 3 * C * E;                 // typed linear gradients
    const long long n2 = (long long)2 * 4 * p.heads * dh * dh;     // relation tensor gradients (att, msg)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n1 + n2; i += (long long)gridDim.x * blockDim.x) {
        if (i < n1) {
            const int col = (int)(i % E);
            const int o = (int)((i / E) % C);
            const int which = (int)((i / ((long long)E * C)) % 3);  // 0 q, 1 k, 2 v
            const int tj = (int)(i / ((long long)E * C * 3));
            const int h = o / dh, t = o - h * dh;
            float val = 0.f;
            if (which == 0) {
                val = dWf(tj, o, col);
            } else {
                for (int ti = 0; ti < 2; ++ti) {
                    const float* R = (which == 1 ? p.rel_att : p.rel_msg) + ((long long)(ti * 2 + tj) * p.heads + h) * dh * dh;
                    const int row0 = (which == 1 ? 1 : 3) * C + ti * C + h * dh;
                    for (int a = 0; a < dh; ++a) {
                        const float r = which == 1 ? R[a * dh + t] : R[t * dh + a];  // k' = A k ; v' = M^T v
                        val = fmaf(r, dWf(tj, row0 + a, col), val);
                    }
                }
            }
            float* W = which == 0 ? p.dqw[tj] : which == 1 ? p.dkw[tj] : p.dvw[tj];
            float* Bv = which == 0 ? p.dqb[tj] : which == 1 ? p.dkb[tj] : p.dvb[tj];
            if (col < C) W[(long long)o * C + col] = val;
            else Bv[o] = val;
        } else {
            long long r = i - n1;
            const int y = (int)(r % dh);
            const int x = (int)((r / dh) % dh);
            const int h = (int)((r / (dh * dh)) % p.heads);
            const int e = (int)((r / ((long long)dh * dh * p.heads)) % 4);
            const int is_msg = (int)(r / ((long long)dh * dh * p.heads * 4));
            const int ti = e >> 1, tj = e & 1;
            // att: dA[x][y] = sum_col dk'[(h,x)][col] * Wk_ext[(h,y)][col] ; msg: dM[x][y] = sum_col Wv_ext[(h,x)][col] * dv'[(h,y)][col]
            const int grow = (is_msg ? 3 : 1) * C + ti * C + h * dh + (is_msg ? y : x);
            const int wrow = h * dh + (is_msg ? x : y);
            const float* Wm = is_msg ? p.vw[tj] : p.kw[tj];
            const float* Bv = is_msg ? p.vb[tj] : p.kb[tj];
            float s = 0.f;
            for (int col = 0; col < C; ++col) s = fmaf(dWf(tj, grow, col), Wm[(long long)wrow * C + col], s);
            s = fmaf(dWf(tj, grow, C), Bv[wrow], s);
            (is_msg ? p.drel_msg : p.drel_att)[((long long)e * p.heads + h) * dh * dh + x * dh + y] = s;
        }
    }
}

// ---- split attention backward
// dw[a][r][c] = sum_p dx[a][p][c] * w_r[a][p][c]
__global__ void __launch_bounds__(256) split_bwd_reduce_kernel(const float* __restrict__ dx, const float* __restrict__ w0,
                                                               const float* __restrict__ w1, const float* __restrict__ w2,
                                                               long long pix, int C, int chunks, float* __restrict__ dw) {
    const int q = C >> 2;
    const int cq = threadIdx.x % q, prow = threadIdx.x / q, ppb = blockDim.x / q;
    const int a = blockIdx.y;
    const long long per = (pix + chunks - 1) / chunks;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(pix, p0 + per);
    float4 s[3];
    for (int r = 0; r < 3; ++r) s[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prow < ppb) {
        for (long long p = p0 + prow; p < p1; p += ppb) {
            const long long o = ((long long)a * pix + p) * C + cq * 4;
            const float4 d = *reinterpret_cast<const float4*>(dx + o);
            const float4 x0 = *reinterpret_cast<const float4*>(w0 + o), x1 = *reinterpret_cast<const float4*>(w1 + o),
                         x2 = *reinterpret_cast<const float4*>(w2 + o);
            s[0].x += d.x * x0.x; s[0].y += d.y * x0.y; s[0].z += d.z * x0.z; s[0].w += d.w * x0.w;
            s[1].x += d.x * x1.x; s[1].y += d.y * x1.y; s[1].z += d.z * x1.z; s[1].w += d.w * x1.w;
            s[2].x += d.x * x2.x; s[2].y += d.y * x2.y; s[2].z += d.z * x2.z; s[2].w += d.w * x2.w;
        }
        for (int r = 0; r < 3; ++r) {
            float* o = dw + ((long long)a * 3 + r) * C + cq * 4;
            atomicAdd(o, s[r].x); atomicAdd(o + 1, s[r].y); atomicAdd(o + 2, s[r].z); atomicAdd(o + 3, s[r].w);
        }
    }
}

// per agent (one block): recompute the tiny MLP, back-propagate dw -> d_gap[a][:] (already divided by pix), and the
// gradients of fc1 / bn1 / fc2 (atomics across agents; caller zeroes them)
__global__ void split_bwd_mlp_kernel(const float* __restrict__ sums, float inv_pix, const float* __restrict__ fc1,
                                     const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                     const float* __restrict__ fc2, const float* __restrict__ dw, int C,
                                     float* __restrict__ d_gap, float* __restrict__ dfc1, float* __restrict__ dln_g,
                                     float* __restrict__ dln_b, float* __restrict__ dfc2) {
    extern __shared__ float sm[];
    float* g = sm;            // [C]   gap
    float* h1 = g + C;        // [C]   fc1 g
    float* xh = h1 + C;       // [C]   xhat
    float* act = xh + C;      // [C]   relu(LN)
    float* da = act + C;      // [3C]  gradient w.r.t. fc2 output
    float* dact = da + 3 * C; // [C]
    float* dh1 = dact + C;    // [C]
    float* red = dh1 + C;     // [4]
    const int a = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) g[c] = sums[(long long)a * C + c] * inv_pix;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < C; ++k) s = fmaf(fc1[(long long)c * C + k], g[k], s);
        h1[c] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int c = 0; c < C; ++c) m += h1[c];
        m /= C;
        float v = 0.f;
        for (int c = 0; c < C; ++c) v += (h1[c] - m) * (h1[c] - m);
        red[0] = m;
        red[1] = rsqrtf(v / C + 1e-5f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        xh[c] = (h1[c] - red[0]) * red[1];
        act[c] = fmaxf(xh[c] * ln_g[c] + ln_b[c], 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s[3];
        for (int r = 0; r < 3; ++r) {
            float t = 0.f;
            const float* w = fc2 + (long long)(r * C + c) * C;
            for (int k = 0; k < C; ++k) t = fmaf(w[k], act[k], t);
            s[r] = t;
        }
        const float mx = fmaxf(s[0], fmaxf(s[1], s[2]));
        float e[3] = {expf(s[0] - mx), expf(s[1] - mx), expf(s[2] - mx)};
        const float inv = 1.f / (e[0] + e[1] + e[2]);
        float dot = 0.f;
        for (int r = 0; r < 3; ++r) {
            e[r] *= inv;
            dot = fmaf(e[r], dw[((long long)a * 3 + r) * C + c], dot);
        }
        for (int r = 0; r < 3; ++r) da[r * C + c] = e[r] * (dw[((long long)a * 3 + r) * C + c] - dot);  // softmax over the radix
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float s = 0.f;
        for (int rc = 0; rc < 3 * C; ++rc) s = fmaf(fc2[(long long)rc * C + k], da[rc], s);
        dact[k] = (xh[k] * ln_g[k] + ln_b[k] > 0.f) ? s : 0.f;  // ReLU gate
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < (long long)3 * C * C; i += blockDim.x)
        atomicAdd(&dfc2[i], da[i / C] * act[i % C]);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&dln_g[c], dact[c] * xh[c]);
        atomicAdd(&dln_b[c], dact[c]);
    }
    if (threadIdx.x == 0) {
        float mg = 0.f, mgx = 0.f;
        for (int c = 0; c < C; ++c) {
            const float gg = dact[c] * ln_g[c];
            mg += gg;
            mgx += gg * xh[c];
        }
        red[2] = mg / C;
        red[3] = mgx / C;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) dh1[c] = red[1] * (dact[c] * ln_g[c] - red[2] - xh[c] * red[3]);
    __syncthreads();
    for (long long i = threadIdx.x; i < (long long)C * C; i += blockDim.x) atomicAdd(&dfc1[i], dh1[i / C] * g[i % C]);
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(fc1[(long long)c * C + k], dh1[c], s);
        d_gap[(long long)a * C + k] = s * inv_pix;
    }
}

// d_win_r[a][p][c] = dx[a][p][c] * wts[a][r][c] + d_gap[a][c]   -> three split operands
__global__ void __launch_bounds__(256) split_bwd_apply_kernel(const float* __restrict__ dx, const float* __restrict__ wts,
                                                              const float* __restrict__ d_gap, long long pix, int C,
                                                              SplitOut o0, SplitOut o1, SplitOut o2, long long total4) {
    const int q = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % q) * 4;
        const long long a = i / (q * pix);
        const float4 d = reinterpret_cast<const float4*>(dx)[i];
        const float4 gg = *reinterpret_cast<const float4*>(d_gap + a * C + c);
        const float* wa = wts + a * 3 * C + c;
        const float4 a0 = *reinterpret_cast<const float4*>(wa), a1 = *reinterpret_cast<const float4*>(wa + C),
                     a2 = *reinterpret_cast<const float4*>(wa + 2 * C);
        store_split4(o0, 4 * i, make_float4(d.x * a0.x + gg.x, d.y * a0.y + gg.y, d.z * a0.z + gg.z, d.w * a0.w + gg.w));
        store_split4(o1, 4 * i, make_float4(d.x * a1.x + gg.x, d.y * a1.y + gg.y, d.z * a1.z + gg.z, d.w * a1.w + gg.w));
        store_split4(o2, 4 * i, make_float4(d.x * a2.x + gg.x, d.y * a2.y + gg.y, d.z * a2.z + gg.z, d.w * a2.w + gg.w));
    }
}

// ---- RTE backward: dvec[a][c] = sum_p dx[a][p][c] (computed by the caller with a2x_channel_stats per agent), then
// dlin_w[c][k] += dvec[a][c] * emb[idx[a]][k], dlin_b[c] += dvec[a][c], demb[idx[a]][k] += sum_c W[c][k] dvec[a][c]
__global__ void rte_bwd_kernel(const double* __restrict__ dvec_sums /* [n][2C], first C used */, const float* __restrict__ emb,
                               const int* __restrict__ idx, const float* __restrict__ W, int C, float* __restrict__ dW,
                               float* __restrict__ db, float* __restrict__ demb) {
    const int a = blockIdx.x;
    const double* dv = dvec_sums + (long long)a * 2 * C;
    const float* e = emb + (long long)idx[a] * C;
    for (long long i = threadIdx.x; i < (long long)C * C; i += blockDim.x) atomicAdd(&dW[i], (float)dv[i / C] * e[i % C]);
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&db[c], (float)dv[c]);
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(W[(long long)c * C + k], (float)dv[c], s);
        atomicAdd(&demb[(long long)idx[a] * C + k], s);
    }
}

}  // namespace a2x

extern "C" {

int a2x_hgt_attention_bwd(const float* qkv, const int* types_dev, const float* key_mask, const float* dout, int n_agents,
                          long long pix, int heads, int dim_head, float scale, float* dqkv, a2x_stream_t stream) {
    A2X_REQUIRE(qkv && types_dev && key_mask && dout && dqkv && n_agents > 0 && n_agents <= a2x::HGT_MAX_AGENTS && pix > 0,
                "hgt_attention_bwd: bad args (at most 12 agents)");
    cudaStream_t st = (cudaStream_t)stream;
    if (dim_head == 16 || dim_head == 32 || dim_head == 64) {
        const long long gpb = 128 / (dim_head / 4), groups = pix * heads;
        const size_t csm = (size_t)gpb * (6 * n_agents * dim_head + 2 * n_agents * n_agents) * sizeof(float);
        if (csm <= 200 * 1024) {
            const unsigned blocks = (unsigned)((groups + gpb - 1) / gpb);
#define A2X_HGTB(DH)                                                                                                  \
    do {                                                                                                              \
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::hgt_attention_bwd_coop_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm)); \
        a2x::hgt_attention_bwd_coop_kernel<DH><<<blocks, 128, csm, st>>>(qkv, types_dev, key_mask, dout, n_agents, pix, heads, scale, dqkv); \
    } while (0)
            if (dim_head == 16) A2X_HGTB(16);
            else if (dim_head == 32) A2X_HGTB(32);
            else A2X_HGTB(64);
#undef A2X_HGTB
            A2X_LAUNCHED();
            A2X_CHECK_CUDA(cudaGetLastError());
            return 0;
        }
    }
    long long b = (pix * heads + 127) / 128;
    if (b > 148 * 16) b = 148 * 16;
    const size_t smem = (size_t)2 * n_agents * n_agents * 128 * sizeof(float);
    A2X_REQUIRE(smem <= 200 * 1024, "hgt_attention_bwd: %d agents do not fit shared memory", n_agents);
    if (dim_head == 32) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::hgt_attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a2x::hgt_attention_bwd_kernel<32><<<(int)b, 128, smem, st>>>(qkv, types_dev, key_mask, dout, n_agents, pix, heads, scale, dqkv);
    } else if (dim_head == 16) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::hgt_attention_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a2x::hgt_attention_bwd_kernel<16><<<(int)b, 128, smem, st>>>(qkv, types_dev, key_mask, dout, n_agents, pix, heads, scale, dqkv);
    }
    else {
        a2x::set_error("hgt_attention_bwd: dim_head %d not in {16, 32}", dim_head);
        return 1;
    }
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_hgt_fold_bwd(const float* dw_fold, const float* db_fold, const float* const* kw, const float* const* kb,
                     const float* const* vw, const float* const* vb, const float* relation_att, const float* relation_msg,
                     int C, int heads, float* const* dqw, float* const* dqb, float* const* dkw, float* const* dkb,
                     float* const* dvw, float* const* dvb, float* drelation_att, float* drelation_msg, a2x_stream_t stream) {
    A2X_REQUIRE(dw_fold && db_fold && kw && kb && vw && vb && relation_att && relation_msg && dqw && dqb && dkw && dkb && dvw &&
                    dvb && drelation_att && drelation_msg && C > 0 && heads > 0 && C % heads == 0,
                "hgt_fold_bwd: bad args");
    a2x::HgtFoldBwdParams p;
    p.dwf = dw_fold; p.dbf = db_fold; p.rel_att = relation_att; p.rel_msg = relation_msg;
    for (int t = 0; t < 2; ++t) {
        p.kw[t] = kw[t]; p.kb[t] = kb[t]; p.vw[t] = vw[t]; p.vb[t] = vb[t];
        p.dqw[t] = dqw[t]; p.dqb[t] = dqb[t]; p.dkw[t] = dkw[t]; p.dkb[t] = dkb[t]; p.dvw[t] = dvw[t]; p.dvb[t] = dvb[t];
    }
    p.drel_att = drelation_att; p.drel_msg = drelation_msg;
    p.C = C; p.heads = heads; p.dh = C / heads;
    const long long total = (long long)2 * 3 * C * (C + 1) + (long long)2 * 4 * heads * p.dh * p.dh;
    a2x::hgt_fold_bwd_kernel<<<a2x::vx_grid(total), 256, 0, (cudaStream_t)stream>>>(p);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_split_attn_bwd(const float* dx, const float* w0, const float* w1, const float* w2, int n_agents, long long pix, int C,
                       const float* fc1, const float* ln_gamma, const float* ln_beta, const float* fc2, const float* sums_saved,
                       const float* weights_saved, float* dw_ws, float* dgap_ws, const a2x_output* d0, const a2x_output* d1,
                       const a2x_output* d2, float* dfc1, float* dln_gamma, float* dln_beta, float* dfc2,
                       a2x_stream_t stream) {
    A2X_REQUIRE(dx && w0 && w1 && w2 && fc1 && ln_gamma && ln_beta && fc2 && sums_saved && weights_saved && dw_ws && dgap_ws &&
                    d0 && d1 && d2 && dfc1 && dln_gamma && dln_beta && dfc2 && n_agents > 0 && pix > 0 && C % 4 == 0 &&
                    256 % (C / 4) == 0,
                "split_attn_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    A2X_CHECK_CUDA(cudaMemsetAsync(dw_ws, 0, (size_t)n_agents * 3 * C * sizeof(float), st));
    const int chunks = 148 * 2 / n_agents > 0 ? 148 * 2 / n_agents : 1;
    a2x::split_bwd_reduce_kernel<<<dim3(chunks, n_agents), 256, 0, st>>>(dx, w0, w1, w2, pix, C, chunks, dw_ws);
    A2X_LAUNCHED();
    a2x::split_bwd_mlp_kernel<<<n_agents, 256, (10 * C + 4) * sizeof(float), st>>>(sums_saved, 1.0f / (float)pix, fc1, ln_gamma,
                                                                                 ln_beta, fc2, dw_ws, C, dgap_ws, dfc1,
                                                                                 dln_gamma, dln_beta, dfc2);
    A2X_LAUNCHED();
    auto so = [](const a2x_output* o) {
        a2x::SplitOut r;
        r.hi = o->hi; r.b16 = (__nv_bfloat16*)o->b16; r.ps = o->b16_plane;
        return r;
    };
    const long long total4 = (long long)n_agents * pix * (C / 4);
    a2x::split_bwd_apply_kernel<<<a2x::vx_grid(total4), 256, 0, st>>>(dx, weights_saved, dgap_ws, pix, C, so(d0), so(d1), so(d2),
                                                                     total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_rte_bwd(const double* dvec_sums, int n_agents, int C, const float* emb_table, const int* emb_idx_dev, const float* lin_w,
                float* dlin_w, float* dlin_b, float* demb_table, a2x_stream_t stream) {
    A2X_REQUIRE(dvec_sums && emb_table && emb_idx_dev && lin_w && dlin_w && dlin_b && demb_table && n_agents > 0 && C > 0,
                "rte_bwd: bad args");
    a2x::rte_bwd_kernel<<<n_agents, 256, 0, (cudaStream_t)stream>>>(dvec_sums, emb_table, emb_idx_dev, lin_w, C, dlin_w, dlin_b,
                                                                   demb_table);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
