// Token-wise kernels of the transformer fusion networks (CoBEVT SwapFusionEncoder, V2X-ViT encoder) around the
// tcgen05 linears (which are the 1x1 tap-GEMM): LayerNorm, 3-D window / grid attention with relative-position bias
// and agent key mask, agent mean + LayerNorm, and regroup (zero-pad every scene to L agents).
// All token tensors are fp32 NHWC: [B*L][H][W][C], so "tokens" are pixels and a window is a strided gather.
//
// Reference semantics:
//   opencood/models/cobevt_modules/base_transformer.py:6-13            PreNormResidual (nn.LayerNorm, eps 1e-5)
//   opencood/models/cobevt_modules/swap_fusion_modules.py:78-127       Attention.forward
//   opencood/models/cobevt_modules/swap_fusion_modules.py:155-195      window "(x w1) (y w2)" / grid "(w1 x) (w2 y)"
//   opencood/models/cobevt_modules/swap_fusion_modules.py:268-275      mean over agents + LayerNorm (+ Linear = GEMM)
//   opencood/models/cobevt_modules/fuse_utils.py:13-63                 regroup
#include "../../include/airv2x_b200.h"
#include "a2x_host.h"
#include "a2x_ptx.cuh"
#include "window_attn_tc.cuh"

namespace a2x {

static __host__ SplitOut tr_split(const a2x_output* o) {
    SplitOut r;
    r.hi = o->hi;
    r.b16 = (__nv_bfloat16*)o->b16;
    r.ps = o->b16_plane;
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; C = 32 * 4 * V channels (V float4 per lane). Two-pass statistics in registers like
// torch.nn.functional.layer_norm (mean, then biased variance of the centred values).
template <int V>
__device__ __forceinline__ void ln_row(float4 (&x)[V], const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float eps, int lane, int C) {
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) s += (x[v].x + x[v].y) + (x[v].z + x[v].w);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        x[v].x -= mean; x[v].y -= mean; x[v].z -= mean; x[v].w -= mean;
        q += (x[v].x * x[v].x + x[v].y * x[v].y) + (x[v].z * x[v].z + x[v].w * x[v].w);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int c = (v * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float4 b = *reinterpret_cast<const float4*>(beta + c);
        x[v].x = x[v].x * rstd * g.x + b.x;
        x[v].y = x[v].y * rstd * g.y + b.y;
        x[v].z = x[v].z * rstd * g.z + b.z;
        x[v].w = x[v].w * rstd * g.w + b.w;
    }
}

template <int V>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int x_cs,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, SplitOut y, int y_cs,
                                                        long long rows) {
    constexpr int C = V * 128;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < rows; r += nwarps) {
        float4 v[V];
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = *reinterpret_cast<const float4*>(x + r * x_cs + (i * 32 + lane) * 4);
        ln_row<V>(v, gamma, beta, eps, lane, C);
#pragma unroll
        for (int i = 0; i < V; ++i) store_split4(y, r * y_cs + (i * 32 + lane) * 4, v[i]);
    }
}

// mean over the L agents of every (scene, pixel) — padded agents included, swap_fusion_modules.py:269-271 — then LN
template <int V>
__global__ void __launch_bounds__(256) agent_mean_ln_kernel(const float* __restrict__ x, int L, long long pix,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, SplitOut y,
                                                            long long rows /* B * pix */) {
    constexpr int C = V * 128;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < rows; r += nwarps) {
        const long long b = r / pix, p = r - b * pix;
        float4 acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) {
            const float* row = x + ((b * L + l) * pix + p) * C;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(row + (i * 32 + lane) * 4);
                acc[i].x += t.x; acc[i].y += t.y; acc[i].z += t.z; acc[i].w += t.w;
            }
        }
        const float inv = 1.f / (float)L;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            acc[i].x *= inv; acc[i].y *= inv; acc[i].z *= inv; acc[i].w *= inv;
        }
        ln_row<V>(acc, gamma, beta, eps, lane, C);
#pragma unroll
        for (int i = 0; i < V; ++i) store_split4(y, r * C + (i * 32 + lane) * 4, acc[i]);
    }
}

// ---------------------------------------------------------------------------------------------- regroup
// dst[b][l] = src[start_b + l] for l < len_b, zeros otherwise; each image is `img` floats (multiple of 4)
__global__ void regroup_kernel(const float* __restrict__ src, const int* __restrict__ scene_start,
                               const int* __restrict__ scene_len, int L, long long img4, SplitOut dst,
                               long long total4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long slot = i / img4, e = i - slot * img4;
        const int b = (int)(slot / L), l = (int)(slot - (long long)b * L);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < scene_len[b]) v = reinterpret_cast<const float4*>(src)[((long long)scene_start[b] + l) * img4 + e];
        store_split4(dst, 4 * i, v);
    }
}

// Agent-parallel form (one agent per GPU): the source images are addressed through a device table of pointers, one per
// agent, which may point into PEER GPUs' memory (NVLink P2P / symmetric memory). The all-gather of the shrunk BEV maps
// is thereby fused into the regroup: every rank pulls each peer's map exactly once, straight into its slot of the
// padded [B*L] token tensor — no intermediate gathered buffer, no separate collective launch.
__global__ void regroup_ptrs_kernel(const float* const* __restrict__ src_ptrs, const int* __restrict__ scene_start,
                                    const int* __restrict__ scene_len, int L, long long img4, SplitOut dst,
                                    long long total4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long slot = i / img4, e = i - slot * img4;
        const int b = (int)(slot / L), l = (int)(slot - (long long)b * L);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < scene_len[b]) {
            const float4* src = reinterpret_cast<const float4*>(src_ptrs[scene_start[b] + l]);
            asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "l"(src + e));  // peer memory: system-scope load, never served from a stale L1 line
        }
        store_split4(dst, 4 * i, v);
    }
}

// ---------------------------------------------------------------------------------------------- window attention
// One CTA per (window, head). Tokens of a window: t = l * w * w + w1 * w + w2  ("(l w1 w2)"), n = L * w * w.
//   window mode: pixel (x*w + w1, y*w + w2);   grid mode: pixel (w1*X + x, w2*Y + y)      (X = H/w, Y = W/w)
// K and V of the window's head live in shared memory; every thread owns QPT query rows (q and the running output in
// registers) and streams over the keys with an online softmax. sim = (q*scale).k + bias[rel(i,j)][head], keys of
// padded agents (key_mask[b][l] == 0) are skipped (= -inf). fp32 throughout.
struct WinAttParams {
    const float* qkv;      // [B*L][H][W][3*D]   (q | k | v, each "(head dim_head)")
    const float* bias;     // [(2L-1)(2w-1)^2][heads]
    const int* key_mask;   // [B][L] or null
    SplitOut out;          // [B*L][H][W][D]
    int B, L, H, W, heads, w, grid_mode;
    float scale;
};

template <int DH, int QPT>
__global__ void __launch_bounds__(64) window_attention_kernel(const WinAttParams p) {
    extern __shared__ float sm[];
    const int ww = p.w * p.w;
    const int n = p.L * ww;
    float* sK = sm;                 // [n][DH]
    float* sV = sK + n * DH;        // [n][DH]
    float* sB = sV + n * DH;        // [(2L-1)(2w-1)^2] bias of this head
    int* sTok = reinterpret_cast<int*>(sB + (2 * p.L - 1) * (2 * p.w - 1) * (2 * p.w - 1));  // [n] pixel-row index
    const int D = p.heads * DH;
    const int X = p.H / p.w, Y = p.W / p.w;
    const int head = blockIdx.x % p.heads;
    int win = blockIdx.x / p.heads;
    const int y = win % Y;
    win /= Y;
    const int x = win % X;
    const int b = win / X;

    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t / ww, r = t - l * ww;
        const int w1 = r / p.w, w2 = r - w1 * p.w;
        const int ph = p.grid_mode ? w1 * X + x : x * p.w + w1;
        const int pw = p.grid_mode ? w2 * Y + y : y * p.w + w2;
        sTok[t] = ((b * p.L + l) * p.H + ph) * p.W + pw;
    }
    const int nb = (2 * p.L - 1) * (2 * p.w - 1) * (2 * p.w - 1);
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sB[i] = p.bias[i * p.heads + head];
    __syncthreads();
    for (int i = threadIdx.x; i < n * (DH / 4); i += blockDim.x) {
        const int t = i / (DH / 4), c = (i - t * (DH / 4)) * 4;
        const float* row = p.qkv + (long long)sTok[t] * (3 * D) + head * DH + c;
        *reinterpret_cast<float4*>(sK + t * DH + c) = *reinterpret_cast<const float4*>(row + D);
        *reinterpret_cast<float4*>(sV + t * DH + c) = *reinterpret_cast<const float4*>(row + 2 * D);
    }
    __syncthreads();

    const int s2 = 2 * p.w - 1;
    for (int q0 = threadIdx.x * QPT; q0 < n; q0 += blockDim.x * QPT) {
        float q[QPT][DH], acc[QPT][DH], m[QPT], den[QPT];
        int ql[QPT], q1[QPT], q2[QPT];
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            const int t = min(q0 + u, n - 1);
            const float* row = p.qkv + (long long)sTok[t] * (3 * D) + head * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(row + c);
                q[u][c] = v.x * p.scale; q[u][c + 1] = v.y * p.scale; q[u][c + 2] = v.z * p.scale; q[u][c + 3] = v.w * p.scale;
            }
#pragma unroll
            for (int c = 0; c < DH; ++c) acc[u][c] = 0.f;
            m[u] = -INFINITY;
            den[u] = 0.f;
            ql[u] = t / ww;
            const int r = t - ql[u] * ww;
            q1[u] = r / p.w;
            q2[u] = r - q1[u] * p.w;
        }
        for (int lj = 0; lj < p.L; ++lj) {
            if (p.key_mask != nullptr && p.key_mask[b * p.L + lj] == 0) continue;  // uniform over the CTA
            for (int r = 0; r < ww; ++r) {
                const int j = lj * ww + r;
                const int k1 = r / p.w, k2 = r - k1 * p.w;
                float kk[DH];
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(sK + j * DH + c);
                    kk[c] = v.x; kk[c + 1] = v.y; kk[c + 2] = v.z; kk[c + 3] = v.w;
                }
                float s[QPT];
#pragma unroll
                for (int u = 0; u < QPT; ++u) {
                    float d0 = 0.f, d1 = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; c += 2) {
                        d0 = fmaf(q[u][c], kk[c], d0);
                        d1 = fmaf(q[u][c + 1], kk[c + 1], d1);
                    }
                    const int bi = ((ql[u] - lj + p.L - 1) * s2 + (q1[u] - k1 + p.w - 1)) * s2 + (q2[u] - k2 + p.w - 1);
                    s[u] = d0 + d1 + sB[bi];
                }
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(sV + j * DH + c);
                    kk[c] = v.x; kk[c + 1] = v.y; kk[c + 2] = v.z; kk[c + 3] = v.w;
                }
#pragma unroll
                for (int u = 0; u < QPT; ++u) {
                    const float mn = fmaxf(m[u], s[u]);
                    const float corr = __expf(m[u] - mn);  // exp(-inf) = 0 on the first key
                    const float e = __expf(s[u] - mn);
                    den[u] = den[u] * corr + e;
#pragma unroll
                    for (int c = 0; c < DH; ++c) acc[u][c] = fmaf(acc[u][c], corr, e * kk[c]);
                    m[u] = mn;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            if (q0 + u >= n) break;
            const float inv = 1.f / den[u];
            const long long off = (long long)sTok[q0 + u] * D + head * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 4)
                store_split4(p.out, off + c,
                             make_float4(acc[u][c] * inv, acc[u][c + 1] * inv, acc[u][c + 2] * inv, acc[u][c + 3] * inv));
        }
    }
}

// Small-window variant (n = L*w*w tokens <= 128 and 128 % n == 0, e.g. the 2x2 / 4x4 single-agent windows of the
// V2X-ViT pyramid, mswin.py:23-108): a 128-thread CTA packs G = 128 / n windows of one head, thread t = (window g,
// query t % n); every thread stages the K / V rows of its own token, so global loads are one contiguous row each.
template <int DH>
__global__ void __launch_bounds__(128) window_attention_small_kernel(const WinAttParams p, int num_windows) {
    extern __shared__ float sm[];
    const int ww = p.w * p.w;
    const int n = p.L * ww;
    constexpr int RS = DH + 4;       // padded row stride (floats): spreads the groups over the banks
    float* sK = sm;                  // [128][RS]
    float* sV = sK + 128 * RS;       // [128][RS]
    float* sB = sV + 128 * RS;
    const int nb = (2 * p.L - 1) * (2 * p.w - 1) * (2 * p.w - 1);
    const int D = p.heads * DH;
    const int X = p.H / p.w, Y = p.W / p.w;
    const int G = 128 / n;
    const int head = blockIdx.x % p.heads;
    const int g = threadIdx.x / n, tq = threadIdx.x - g * n;
    int win = (blockIdx.x / p.heads) * G + g;
    const bool valid = win < num_windows;
    if (!valid) win = num_windows - 1;
    const int y = win % Y;
    const int x = (win / Y) % X;
    const int b = win / (Y * X);
    const int l = tq / ww, r = tq - l * ww;
    const int w1 = r / p.w, w2 = r - w1 * p.w;
    const int ph = p.grid_mode ? w1 * X + x : x * p.w + w1;
    const int pw = p.grid_mode ? w2 * Y + y : y * p.w + w2;
    const long long tok = ((long long)(b * p.L + l) * p.H + ph) * p.W + pw;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sB[i] = p.bias[i * p.heads + head];
    const float* row = p.qkv + tok * (3 * D) + head * DH;
    float q[DH];
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(row + c);
        q[c] = qv.x * p.scale; q[c + 1] = qv.y * p.scale; q[c + 2] = qv.z * p.scale; q[c + 3] = qv.w * p.scale;
        *reinterpret_cast<float4*>(sK + threadIdx.x * RS + c) = *reinterpret_cast<const float4*>(row + D + c);
        *reinterpret_cast<float4*>(sV + threadIdx.x * RS + c) = *reinterpret_cast<const float4*>(row + 2 * D + c);
    }
    __syncthreads();
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    float mx = -INFINITY, den = 0.f;
    const int s2 = 2 * p.w - 1;
    for (int j = 0; j < n; ++j) {
        const int lj = j / ww, rj = j - lj * ww;
        if (p.key_mask != nullptr && p.key_mask[b * p.L + lj] == 0) continue;
        const int k1 = rj / p.w, k2 = rj - k1 * p.w;
        const float* kr = sK + (g * n + j) * RS;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(kr + c);
            d0 = fmaf(q[c], kv.x, d0); d1 = fmaf(q[c + 1], kv.y, d1);
            d0 = fmaf(q[c + 2], kv.z, d0); d1 = fmaf(q[c + 3], kv.w, d1);
        }
        const float sc = d0 + d1 + sB[((l - lj + p.L - 1) * s2 + (w1 - k1 + p.w - 1)) * s2 + (w2 - k2 + p.w - 1)];
        const float mn = fmaxf(mx, sc);
        const float corr = __expf(mx - mn), e = __expf(sc - mn);
        den = den * corr + e;
        const float* vr = sV + (g * n + j) * RS;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(vr + c);
            acc[c] = fmaf(acc[c], corr, e * vv.x); acc[c + 1] = fmaf(acc[c + 1], corr, e * vv.y);
            acc[c + 2] = fmaf(acc[c + 2], corr, e * vv.z); acc[c + 3] = fmaf(acc[c + 3], corr, e * vv.w);
        }
        mx = mn;
    }
    if (valid) {
        const float inv = 1.f / den;
        const long long off = tok * D + head * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4)
            store_split4(p.out, off + c, make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv));
    }
}

// tcgen05 window attention (window_attn_tc.cuh) applies to 4x4 windows, 32-wide heads, 32..128 tokens in multiples of 16,
// at most 8 agents (key-mask bits); A2X_ATTN_SIMT=1 forces the fp32 CUDA-core kernels (A/B debugging)
static bool wt_eligible(int window, int dim_head, int n, int L) {
    static int simt = -1;
    if (simt < 0) {
        const char* e = getenv("A2X_ATTN_SIMT");
        simt = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return !simt && window == WT_W && dim_head == WT_DH && n % 16 == 0 && n >= 32 && n <= 128 && L <= 8;
}
// persistent grid: a multiple of `heads` CTAs (a CTA keeps one head), ctas_per_sm per SM, never more than the work
static int wt_grid(int num_windows, int heads, int ctas_per_sm) {
    int groups = (148 * ctas_per_sm) / heads;
    if (groups > num_windows) groups = num_windows;
    if (groups < 1) groups = 1;
    return groups * heads;
}

static int row_grid(long long rows) {
    long long b = (rows + 7) / 8;  // 8 warps (rows) per 256-thread block
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace a2x

using namespace a2x;

extern "C" {

int a2x_layernorm_fwd(const float* x, int x_cs, long long rows, int C, const float* gamma, const float* beta, float eps,
                      const a2x_output* y, a2x_stream_t stream) {
    A2X_REQUIRE(x && gamma && beta && y && y->hi && rows > 0 && x_cs >= C && y->cs >= C, "layernorm_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = row_grid(rows);
    if (C == 128) layernorm_kernel<1><<<g, 256, 0, st>>>(x, x_cs, gamma, beta, eps, tr_split(y), y->cs, rows);
    else if (C == 256) layernorm_kernel<2><<<g, 256, 0, st>>>(x, x_cs, gamma, beta, eps, tr_split(y), y->cs, rows);
    else if (C == 384) layernorm_kernel<3><<<g, 256, 0, st>>>(x, x_cs, gamma, beta, eps, tr_split(y), y->cs, rows);
    else if (C == 512) layernorm_kernel<4><<<g, 256, 0, st>>>(x, x_cs, gamma, beta, eps, tr_split(y), y->cs, rows);
    else {
        set_error("layernorm_fwd: channel count %d not in {128, 256, 384, 512}", C);
        return 1;
    }
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_agent_mean_layernorm(const float* x, int B, int L, long long pix, int C, const float* gamma, const float* beta,
                             float eps, const a2x_output* y, a2x_stream_t stream) {
    A2X_REQUIRE(x && gamma && beta && y && y->hi && B > 0 && L > 0 && pix > 0 && y->cs == C,
                "agent_mean_layernorm: bad args (dense output expected)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * pix;
    const int g = row_grid(rows);
    if (C == 128) agent_mean_ln_kernel<1><<<g, 256, 0, st>>>(x, L, pix, gamma, beta, eps, tr_split(y), rows);
    else if (C == 256) agent_mean_ln_kernel<2><<<g, 256, 0, st>>>(x, L, pix, gamma, beta, eps, tr_split(y), rows);
    else if (C == 384) agent_mean_ln_kernel<3><<<g, 256, 0, st>>>(x, L, pix, gamma, beta, eps, tr_split(y), rows);
    else if (C == 512) agent_mean_ln_kernel<4><<<g, 256, 0, st>>>(x, L, pix, gamma, beta, eps, tr_split(y), rows);
    else {
        set_error("agent_mean_layernorm: channel count %d not in {128, 256, 384, 512}", C);
        return 1;
    }
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_regroup(const float* src, const int* scene_start, const int* scene_len, int B, int L, long long img_elems,
                const a2x_output* dst, a2x_stream_t stream) {
    A2X_REQUIRE(src && scene_start && scene_len && dst && dst->hi && B > 0 && L > 0 && img_elems > 0 && img_elems % 4 == 0,
                "regroup: bad args (image size must be a multiple of 4 floats)");
    const long long total4 = (long long)B * L * (img_elems / 4);
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    regroup_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, scene_start, scene_len, L, img_elems / 4,
                                                                 tr_split(dst), total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_regroup_ptrs(const float* const* src_ptrs_dev, const int* scene_start, const int* scene_len, int B, int L,
                     long long img_elems, const a2x_output* dst, a2x_stream_t stream) {
    A2X_REQUIRE(src_ptrs_dev && scene_start && scene_len && dst && dst->hi && B > 0 && L > 0 && img_elems > 0 &&
                    img_elems % 4 == 0,
                "regroup_ptrs: bad args (image size must be a multiple of 4 floats)");
    const long long total4 = (long long)B * L * (img_elems / 4);
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    regroup_ptrs_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src_ptrs_dev, scene_start, scene_len, L,
                                                                      img_elems / 4, tr_split(dst), total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_window_attention_fwd(const float* qkv, const float* bias_table, const int* key_mask, int B, int L, int H, int W,
                             int heads, int dim_head, int window, int grid_mode, float scale, const a2x_output* out,
                             a2x_stream_t stream) {
    A2X_REQUIRE(qkv && bias_table && out && out->hi && B > 0 && L > 0 && heads > 0 && window > 0,
                "window_attention_fwd: bad args");
    A2X_REQUIRE(H % window == 0 && W % window == 0, "window_attention_fwd: H, W must be multiples of the window");
    A2X_REQUIRE(out->cs == heads * dim_head, "window_attention_fwd: dense [.., heads*dim_head] output expected");
    WinAttParams p;
    p.qkv = qkv; p.bias = bias_table; p.key_mask = key_mask; p.out = tr_split(out);
    p.B = B; p.L = L; p.H = H; p.W = W; p.heads = heads; p.w = window; p.grid_mode = grid_mode; p.scale = scale;
    const int n = L * window * window;
    const int nb = (2 * L - 1) * (2 * window - 1) * (2 * window - 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 64 && 128 % n == 0) {  // small windows: pack 128 / n windows per CTA
        const int num_windows = B * (H / window) * (W / window);
        const int G = 128 / n;
        const size_t sm_small = (size_t)(2 * 128 * (dim_head + 4) + nb) * sizeof(float);
        const long long gsm = (long long)((num_windows + G - 1) / G) * heads;
#define A2X_WAS(DH)                                                                                             \
    do {                                                                                                        \
        A2X_CHECK_CUDA(cudaFuncSetAttribute(window_attention_small_kernel<DH>,                                  \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));       \
        window_attention_small_kernel<DH><<<(unsigned)gsm, 128, sm_small, st>>>(p, num_windows);                \
    } while (0)
        if (dim_head == 16) A2X_WAS(16);
        else if (dim_head == 32) A2X_WAS(32);
        else if (dim_head == 64) A2X_WAS(64);
        else {
            set_error("window_attention_fwd: dim_head %d not in {16, 32, 64}", dim_head);
            return 1;
        }
#undef A2X_WAS
        A2X_LAUNCHED();
        A2X_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (wt_eligible(window, dim_head, n, L)) {  // tcgen05 path: S = QK^T and O = PV as bf16x3 UMMA tiles, softmax on the TMEM row
        WinTcParams q;
        q.qkv = qkv; q.dout = nullptr; q.bias = bias_table; q.key_mask = key_mask; q.out = tr_split(out); q.dbias = nullptr;
        q.B = B; q.L = L; q.H = H; q.W = W; q.heads = heads; q.grid_mode = grid_mode; q.scale = scale;
        const int num_windows = B * (H / window) * (W / window);
        const int smem_tc = wt_smem_bytes(n, false);
        A2X_CHECK_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc));
        window_attention_tc_kernel<false><<<wt_grid(num_windows, heads, 1), WT_THREADS, smem_tc, st>>>(q, num_windows);
        A2X_LAUNCHED();
        A2X_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = (size_t)(2 * n * dim_head + nb + n) * sizeof(float);
    A2X_REQUIRE(smem <= 200 * 1024, "window_attention_fwd: window of %d tokens does not fit shared memory", n);
    const long long grid = (long long)B * (H / window) * (W / window) * heads;
#define A2X_WA(DH, QPT)                                                                                         \
    do {                                                                                                        \
        A2X_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<DH, QPT>,                                   \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        window_attention_kernel<DH, QPT><<<(unsigned)grid, 64, smem, st>>>(p);                                  \
    } while (0)
    if (dim_head == 16) A2X_WA(16, 2);
    else if (dim_head == 32) A2X_WA(32, 2);
    else if (dim_head == 64) A2X_WA(64, 1);
    else {
        set_error("window_attention_fwd: dim_head %d not in {16, 32, 64}", dim_head);
        return 1;
    }
#undef A2X_WA
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

// =====================================================================================================================
// Backward kernels of the transformer fusion (training step of the CoBEVT path): LayerNorm, GELU, window attention.
// Autograd of nn.LayerNorm / nn.GELU / Attention.forward (cobevt_modules/base_transformer.py:6-28,
// swap_fusion_modules.py:78-127); everything is recomputed from the saved inputs, nothing but x / qkv is stored.
namespace a2x {

// dx_accum[r][:] += LN'(x[r]) applied to dy[r];  dgamma[c] += sum_r dy*xhat, dbeta[c] += sum_r dy  (double accumulators)
template <int V>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, int x_cs,
                                                            const float* __restrict__ dy, int dy_cs,
                                                            const float* __restrict__ gamma, float eps,
                                                            float* __restrict__ dx_accum, int dx_cs, long long rows,
                                                            double* __restrict__ dgamma, double* __restrict__ dbeta) {
    constexpr int C = V * 128;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float4 gq[V], ag[V], ab[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        gq[i] = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
        ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long r = warp0; r < rows; r += nwarps) {
        float4 xv[V], dv[V];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            xv[i] = *reinterpret_cast<const float4*>(x + r * x_cs + (i * 32 + lane) * 4);
            dv[i] = *reinterpret_cast<const float4*>(dy + r * dy_cs + (i * 32 + lane) * 4);
            s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
            q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
            ag[i].x += dv[i].x * xv[i].x; ag[i].y += dv[i].y * xv[i].y; ag[i].z += dv[i].z * xv[i].z; ag[i].w += dv[i].w * xv[i].w;
            ab[i].x += dv[i].x; ab[i].y += dv[i].y; ab[i].z += dv[i].z; ab[i].w += dv[i].w;
            dv[i].x *= gq[i].x; dv[i].y *= gq[i].y; dv[i].z *= gq[i].z; dv[i].w *= gq[i].w;  // g = dy * gamma
            sg += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
            sgx += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
        }
        const float mg = warp_sum(sg) / (float)C, mgx = warp_sum(sgx) / (float)C;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float4* o = reinterpret_cast<float4*>(dx_accum + r * dx_cs + (i * 32 + lane) * 4);
            float4 t = *o;
            t.x += rstd * (dv[i].x - mg - xv[i].x * mgx);
            t.y += rstd * (dv[i].y - mg - xv[i].y * mgx);
            t.z += rstd * (dv[i].z - mg - xv[i].z * mgx);
            t.w += rstd * (dv[i].w - mg - xv[i].w * mgx);
            *o = t;
        }
    }
    // block combine of the per-warp column sums, then one double atomic per column per block
    __shared__ float red[8][2][V * 128];
    const int wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        *reinterpret_cast<float4*>(&red[wid][0][(i * 32 + lane) * 4]) = ag[i];
        *reinterpret_cast<float4*>(&red[wid][1][(i * 32 + lane) * 4]) = ab[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float tg = 0.f, tb = 0.f;
        for (int w = 0; w < 8; ++w) {
            tg += red[w][0][c];
            tb += red[w][1][c];
        }
        atomicAdd(&dgamma[c], (double)tg);
        atomicAdd(&dbeta[c], (double)tb);
    }
}

// y = gelu(x) (erf form) as a split operand;  dx = dy * gelu'(x) as a split operand
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, long long n4, SplitOut y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = 0.5f * v.x * (1.f + erff(v.x * 0.70710678118654752f));
        v.y = 0.5f * v.y * (1.f + erff(v.y * 0.70710678118654752f));
        v.z = 0.5f * v.z * (1.f + erff(v.z * 0.70710678118654752f));
        v.w = 0.5f * v.w * (1.f + erff(v.w * 0.70710678118654752f));
        store_split4(y, 4 * i, v);
    }
}
__device__ __forceinline__ float gelu_grad(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                       long long n4, SplitOut dx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 d = reinterpret_cast<const float4*>(dy)[i];
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        store_split4(dx, 4 * i, make_float4(d.x * gelu_grad(v.x), d.y * gelu_grad(v.y), d.z * gelu_grad(v.z), d.w * gelu_grad(v.w)));
    }
}

// ---- window attention backward. One CTA per (window, head), 128 threads, Q (pre-scaled) / K / V / dO in shared memory.
// phase 1 (thread = query i): softmax statistics m_i, l_i, D_i = sum_j P_ij dP_ij, and dQ_i;
// phase 2 (thread = key j):   dK_j, dV_j, and the relative-position-bias gradient (shared-memory table, flushed by atomics).
struct WinAttBwdParams {
    const float* qkv;      // [B*L][H][W][3*D]
    const float* dout;     // [B*L][H][W][D]   gradient w.r.t. the attention output (before to_out)
    const float* bias;     // [(2L-1)(2w-1)^2][heads]
    const int* key_mask;   // [B][L] or null
    SplitOut dqkv;         // [B*L][H][W][3*D]  (written: every token belongs to exactly one window per call); fp32 and / or split planes
    float* dbias;          // [(2L-1)(2w-1)^2][heads], accumulated with atomics
    int B, L, H, W, heads, w, grid_mode;
    float scale;
};

template <int DH>
__global__ void __launch_bounds__(128) window_attention_bwd_kernel(const WinAttBwdParams p) {
    // The n x n probability matrix of the (window, head) lives in shared memory (n <= 128: 64 KB), so every score is
    // computed once:   P = softmax(S)            (thread = query i, one sweep over the keys)
    //                  dV_j = sum_i P_ij dO_i    (thread = key j)
    //                  O_i = sum_j P_ij V_j, D_i = dO_i . O_i,  dS_ij = P_ij (dO_i . V_j - D_i) overwrites P_ij, dQ_i
    //                  dK_j = sum_i dS_ij Q_i, bias gradient      (thread = key j)
    extern __shared__ float sm[];
    const int ww = p.w * p.w;
    const int n = p.L * ww;
    constexpr int RS = DH + 4;      // float4-aligned rows, staggered banks
    float* sQ = sm;                 // [n][RS]  q * scale
    float* sK = sQ + n * RS;
    float* sV = sK + n * RS;
    float* sO = sV + n * RS;        // dO
    float* sP = sO + n * RS;        // [n][n + 1]
    const int PS = n + 1;
    const int nb = (2 * p.L - 1) * (2 * p.w - 1) * (2 * p.w - 1);
    float* sB = sP + n * PS;        // [nb] bias of this head
    float* sdB = sB + nb;           // [nb] bias gradient of this CTA
    int* sTok = reinterpret_cast<int*>(sdB + nb);
    const int D = p.heads * DH;
    const int X = p.H / p.w, Y = p.W / p.w;
    const int head = blockIdx.x % p.heads;
    int win = blockIdx.x / p.heads;
    const int y = win % Y;
    win /= Y;
    const int x = win % X;
    const int b = win / X;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t / ww, r = t - l * ww;
        const int w1 = r / p.w, w2 = r - w1 * p.w;
        const int ph = p.grid_mode ? w1 * X + x : x * p.w + w1;
        const int pw = p.grid_mode ? w2 * Y + y : y * p.w + w2;
        sTok[t] = ((b * p.L + l) * p.H + ph) * p.W + pw;
    }
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        sB[i] = p.bias[i * p.heads + head];
        sdB[i] = 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * (DH / 4); i += blockDim.x) {
        const int t = i / (DH / 4), c = (i - t * (DH / 4)) * 4;
        const float* row = p.qkv + (long long)sTok[t] * (3 * D) + head * DH + c;
        float4 q4 = *reinterpret_cast<const float4*>(row);
        q4.x *= p.scale; q4.y *= p.scale; q4.z *= p.scale; q4.w *= p.scale;
        *reinterpret_cast<float4*>(sQ + t * RS + c) = q4;
        *reinterpret_cast<float4*>(sK + t * RS + c) = *reinterpret_cast<const float4*>(row + D);
        *reinterpret_cast<float4*>(sV + t * RS + c) = *reinterpret_cast<const float4*>(row + 2 * D);
        *reinterpret_cast<float4*>(sO + t * RS + c) =
            *reinterpret_cast<const float4*>(p.dout + (long long)sTok[t] * D + head * DH + c);
    }
    __syncthreads();
    const int s2 = 2 * p.w - 1;
    auto valid_key = [&](int lj) { return p.key_mask == nullptr || p.key_mask[b * p.L + lj] != 0; };
    auto ld = [&](const float* base, int row, float (&v)[DH]) {
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(base + row * RS + c);
            v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
        }
    };
    // ---- P = softmax(S), row i per thread
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float q[DH];
        ld(sQ, i, q);
        const int li = i / ww, ri = i - li * ww, i1 = ri / p.w, i2 = ri - i1 * p.w;
        float m = -INFINITY;
        for (int lj = 0; lj < p.L; ++lj) {
            const bool ok = valid_key(lj);
            for (int r = 0; r < ww; ++r) {
                const int j = lj * ww + r;
                float sc = -INFINITY;
                if (ok) {
                    const int k1 = r / p.w, k2 = r - k1 * p.w;
                    float kk[DH];
                    ld(sK, j, kk);
                    float d0 = 0.f, d1 = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; c += 2) {
                        d0 = fmaf(q[c], kk[c], d0);
                        d1 = fmaf(q[c + 1], kk[c + 1], d1);
                    }
                    sc = d0 + d1 + sB[((li - lj + p.L - 1) * s2 + (i1 - k1 + p.w - 1)) * s2 + (i2 - k2 + p.w - 1)];
                    m = fmaxf(m, sc);
                }
                sP[i * PS + j] = sc;
            }
        }
        float l = 0.f;
        for (int j = 0; j < n; ++j) {
            const float e = expf(sP[i * PS + j] - m);  // exp(-inf) = 0 for masked keys
            sP[i * PS + j] = e;
            l += e;
        }
        const float inv = 1.f / l;
        for (int j = 0; j < n; ++j) sP[i * PS + j] *= inv;
    }
    __syncthreads();
    // ---- dV_j = sum_i P_ij dO_i, column j per thread
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float dv[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) dv[c] = 0.f;
        for (int i = 0; i < n; ++i) {
            const float pij = sP[i * PS + j];
            float o[DH];
            ld(sO, i, o);
#pragma unroll
            for (int c = 0; c < DH; ++c) dv[c] = fmaf(pij, o[c], dv[c]);
        }
        const long long out = (long long)sTok[j] * (3 * D) + 2 * D + head * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4) store_split4(p.dqkv, out + c, make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]));
    }
    __syncthreads();
    // ---- D_i = dO_i . O_i ; dS_ij = P_ij (dO_i . V_j - D_i) -> sP ; dQ_i
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float o[DH], acc[DH];
        ld(sO, i, o);
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = 0.f;
        for (int j = 0; j < n; ++j) {
            const float pij = sP[i * PS + j];
            float v[DH];
            ld(sV, j, v);
#pragma unroll
            for (int c = 0; c < DH; ++c) acc[c] = fmaf(pij, v[c], acc[c]);
        }
        float Dv = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) Dv = fmaf(o[c], acc[c], Dv);
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = 0.f;  // now dQ
        for (int j = 0; j < n; ++j) {
            float v[DH];
            ld(sV, j, v);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c = 0; c < DH; c += 2) {
                d0 = fmaf(o[c], v[c], d0);
                d1 = fmaf(o[c + 1], v[c + 1], d1);
            }
            const float ds = sP[i * PS + j] * (d0 + d1 - Dv);
            sP[i * PS + j] = ds;
            ld(sK, j, v);
#pragma unroll
            for (int c = 0; c < DH; ++c) acc[c] = fmaf(ds, v[c], acc[c]);
        }
        const long long out = (long long)sTok[i] * (3 * D) + head * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4)
            store_split4(p.dqkv, out + c,
                         make_float4(acc[c] * p.scale, acc[c + 1] * p.scale, acc[c + 2] * p.scale, acc[c + 3] * p.scale));
    }
    __syncthreads();
    // ---- dK_j = sum_i dS_ij Q_i (sQ carries the scale) ; bias gradient
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float dk[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) dk[c] = 0.f;
        const int lj = j / ww, rj = j - lj * ww, k1 = rj / p.w, k2 = rj - k1 * p.w;
        for (int i = 0; i < n; ++i) {
            const float ds = sP[i * PS + j];
            float q[DH];
            ld(sQ, i, q);
#pragma unroll
            for (int c = 0; c < DH; ++c) dk[c] = fmaf(ds, q[c], dk[c]);
            if (ds != 0.f) {
                const int li = i / ww, ri = i - li * ww;
                atomicAdd(&sdB[((li - lj + p.L - 1) * s2 + (ri / p.w - k1 + p.w - 1)) * s2 + (ri % p.w - k2 + p.w - 1)], ds);
            }
        }
        const long long out = (long long)sTok[j] * (3 * D) + D + head * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4) store_split4(p.dqkv, out + c, make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (sdB[i] != 0.f) atomicAdd(&p.dbias[i * p.heads + head], sdB[i]);
}

// Small-window backward (NT = L*w*w in {4, 16} tokens): 128 / NT windows per CTA, no atomics on the data path.
// Phase 1, thread = (window g, query i): probability row and dS row in registers -> dQ; both rows go to shared memory.
// Phase 2, the same thread as key j of its window: dK_j = sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i from the shared rows.
// The relative-position-bias gradient is gathered per table entry (one thread per entry), one global atomic per entry and CTA.
template <int DH, int NT>
__global__ void __launch_bounds__(128) window_attention_small_bwd_kernel(const WinAttBwdParams p, int num_windows) {
    extern __shared__ float sm[];
    const int ww = p.w * p.w;
    constexpr int RS = DH + 4, PS = NT + 1, G = 128 / NT;
    float* sK = sm;                   // [128][RS]
    float* sV = sK + 128 * RS;
    float* sQ = sV + 128 * RS;        // scaled q
    float* sG = sQ + 128 * RS;        // dO
    float* sP = sG + 128 * RS;        // [128][PS]
    float* sS = sP + 128 * PS;
    const int nb = (2 * p.L - 1) * (2 * p.w - 1) * (2 * p.w - 1);
    float* sB = sS + 128 * PS;
    const int D = p.heads * DH;
    const int X = p.H / p.w, Y = p.W / p.w;
    const int head = blockIdx.x % p.heads;
    const int g = threadIdx.x / NT, tq = threadIdx.x - g * NT;
    const int win0 = (blockIdx.x / p.heads) * G;
    int win = win0 + g;
    const bool valid = win < num_windows;
    if (!valid) win = num_windows - 1;
    const int y = win % Y;
    const int x = (win / Y) % X;
    const int b = win / (Y * X);
    const int l = tq / ww, r = tq - l * ww;
    const int w1 = r / p.w, w2 = r - w1 * p.w;
    const int ph = p.grid_mode ? w1 * X + x : x * p.w + w1;
    const int pw = p.grid_mode ? w2 * Y + y : y * p.w + w2;
    const long long tok = ((long long)(b * p.L + l) * p.H + ph) * p.W + pw;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) sB[i] = p.bias[i * p.heads + head];
    const float* row = p.qkv + tok * (3 * D) + head * DH;
    const int s2 = 2 * p.w - 1;
    float pr[NT], ds[NT];
    float mx = -INFINITY;
    {
        float q[DH], go[DH];
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            float4 qv = *reinterpret_cast<const float4*>(row + c);
            qv.x *= p.scale; qv.y *= p.scale; qv.z *= p.scale; qv.w *= p.scale;
            q[c] = qv.x; q[c + 1] = qv.y; q[c + 2] = qv.z; q[c + 3] = qv.w;
            *reinterpret_cast<float4*>(sQ + threadIdx.x * RS + c) = qv;
            *reinterpret_cast<float4*>(sK + threadIdx.x * RS + c) = *reinterpret_cast<const float4*>(row + D + c);
            *reinterpret_cast<float4*>(sV + threadIdx.x * RS + c) = *reinterpret_cast<const float4*>(row + 2 * D + c);
            const float4 gv = *reinterpret_cast<const float4*>(p.dout + tok * D + head * DH + c);
            go[c] = gv.x; go[c + 1] = gv.y; go[c + 2] = gv.z; go[c + 3] = gv.w;
            *reinterpret_cast<float4*>(sG + threadIdx.x * RS + c) = gv;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            pr[j] = -INFINITY;
            ds[j] = 0.f;
            const int lj = j / ww, rj = j - lj * ww;
            if (p.key_mask != nullptr && p.key_mask[b * p.L + lj] == 0) continue;
            const float* kr = sK + (g * NT + j) * RS;
            const float* vr = sV + (g * NT + j) * RS;
            float a0 = sB[((l - lj + p.L - 1) * s2 + (w1 - rj / p.w + p.w - 1)) * s2 + (w2 - rj % p.w + p.w - 1)], a1 = 0.f;
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kr + c);
                const float4 v4 = *reinterpret_cast<const float4*>(vr + c);
                a0 = fmaf(q[c], k4.x, a0); a1 = fmaf(q[c + 1], k4.y, a1);
                a0 = fmaf(q[c + 2], k4.z, a0); a1 = fmaf(q[c + 3], k4.w, a1);
                d0 = fmaf(go[c], v4.x, d0); d1 = fmaf(go[c + 1], v4.y, d1);
                d0 = fmaf(go[c + 2], v4.z, d0); d1 = fmaf(go[c + 3], v4.w, d1);
            }
            pr[j] = a0 + a1;
            ds[j] = d0 + d1;
            mx = fmaxf(mx, pr[j]);
        }
    }
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        pr[j] = expf(pr[j] - mx);
        lsum += pr[j];
    }
    const float inv = valid ? 1.f / lsum : 0.f;  // a clamped (out-of-range) window contributes nothing
    float Dv = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        pr[j] *= inv;
        Dv = fmaf(pr[j], ds[j], Dv);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        ds[j] = pr[j] * (ds[j] - Dv);
        sP[threadIdx.x * PS + j] = pr[j];
        sS[threadIdx.x * PS + j] = ds[j];
    }
    if (valid) {
        const long long out = tok * (3 * D) + head * DH;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
            float acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float* kr = sK + (g * NT + j) * RS + c0;
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(kr + c);
                    acc[c] = fmaf(ds[j], k4.x, acc[c]); acc[c + 1] = fmaf(ds[j], k4.y, acc[c + 1]);
                    acc[c + 2] = fmaf(ds[j], k4.z, acc[c + 2]); acc[c + 3] = fmaf(ds[j], k4.w, acc[c + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; c += 4)
                store_split4(p.dqkv, out + c0 + c,
                             make_float4(acc[c] * p.scale, acc[c + 1] * p.scale, acc[c + 2] * p.scale, acc[c + 3] * p.scale));
        }
    }
    __syncthreads();
    // ---- phase 2: this thread as key tq of window g
    if (valid) {
        const long long out = tok * (3 * D) + head * DH;
        float sj[NT], pj[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            sj[i] = sS[(g * NT + i) * PS + tq];
            pj[i] = sP[(g * NT + i) * PS + tq];
        }
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
            float dk[16], dv[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) dk[c] = dv[c] = 0.f;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const float* qr = sQ + (g * NT + i) * RS + c0;
                const float* gr = sG + (g * NT + i) * RS + c0;
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float4 q4 = *reinterpret_cast<const float4*>(qr + c);
                    const float4 g4 = *reinterpret_cast<const float4*>(gr + c);
                    dk[c] = fmaf(sj[i], q4.x, dk[c]); dk[c + 1] = fmaf(sj[i], q4.y, dk[c + 1]);
                    dk[c + 2] = fmaf(sj[i], q4.z, dk[c + 2]); dk[c + 3] = fmaf(sj[i], q4.w, dk[c + 3]);
                    dv[c] = fmaf(pj[i], g4.x, dv[c]); dv[c + 1] = fmaf(pj[i], g4.y, dv[c + 1]);
                    dv[c + 2] = fmaf(pj[i], g4.z, dv[c + 2]); dv[c + 3] = fmaf(pj[i], g4.w, dv[c + 3]);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                store_split4(p.dqkv, out + D + c0 + c, make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]));
                store_split4(p.dqkv, out + 2 * D + c0 + c, make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]));
            }
        }
    }
    // ---- bias gradient: one thread per table entry gathers dS over the CTA's windows
    for (int e = threadIdx.x; e < nb; e += blockDim.x) {
        const int e2 = e % s2 - (p.w - 1), e1 = (e / s2) % s2 - (p.w - 1), el = e / (s2 * s2) - (p.L - 1);
        float acc = 0.f;
        for (int i = 0; i < NT; ++i) {
            const int li = i / ww, ri = i - li * ww, i1 = ri / p.w, i2 = ri - i1 * p.w;
            const int lj = li - el, k1 = i1 - e1, k2 = i2 - e2;
            if (lj < 0 || lj >= p.L || k1 < 0 || k1 >= p.w || k2 < 0 || k2 >= p.w) continue;
            const int j = lj * ww + k1 * p.w + k2;
            for (int gg = 0; gg < G; ++gg) acc += sS[(gg * NT + i) * PS + j];
        }
        if (acc != 0.f) atomicAdd(&p.dbias[e * p.heads + head], acc);
    }
}

// Windows too large for the n x n probability matrix in shared memory (128 < n <= 256 tokens: the 16 x 16 windows of the
// legacy V2X-ViT pyramid, v2xvit_modules/mswin.py:66-82 with window_size [4, 8, 16]). Nothing n x n is stored: the scores
// are recomputed from the rows. Phase A (thread = query i; K, V in shared memory): row maximum, 1 / row sum, D_i =
// sum_j P_ij (dO_i . V_j), then dQ_i. Phase B (thread = key j; scaled Q, dO in shared memory, the row statistics of phase
// A): dV_j = sum_i P_ij dO_i, then dK_j = sum_i dS_ij Q_i and the bias gradient. A fallback for the small legacy maps, not
// a tuned kernel: 3 + 2 sweeps of n dot products per thread.
template <int DH>
__global__ void __launch_bounds__(256) window_attention_bwd_large_kernel(const WinAttBwdParams p) {
    extern __shared__ float sm[];
    const int ww = p.w * p.w;
    const int n = p.L * ww;
    constexpr int RS = DH + 4;
    float* sA = sm;                 // phase A: K       phase B: q * scale
    float* sC = sA + n * RS;        // phase A: V       phase B: dO
    float* sM = sC + n * RS;        // [n] row maximum
    float* sL = sM + n;             // [n] 1 / row sum
    float* sD = sL + n;             // [n] D_i
    const int nb = (2 * p.L - 1) * (2 * p.w - 1) * (2 * p.w - 1);
    float* sB = sD + n;             // [nb] bias of this head
    float* sdB = sB + nb;           // [nb] bias gradient of this CTA
    int* sTok = reinterpret_cast<int*>(sdB + nb);
    const int D = p.heads * DH;
    const int X = p.H / p.w, Y = p.W / p.w;
    const int head = blockIdx.x % p.heads;
    int win = blockIdx.x / p.heads;
    const int y = win % Y;
    win /= Y;
    const int x = win % X;
    const int b = win / X;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int l = t / ww, r = t - l * ww;
        const int w1 = r / p.w, w2 = r - w1 * p.w;
        const int ph = p.grid_mode ? w1 * X + x : x * p.w + w1;
        const int pw = p.grid_mode ? w2 * Y + y : y * p.w + w2;
        sTok[t] = ((b * p.L + l) * p.H + ph) * p.W + pw;
    }
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        sB[i] = p.bias[i * p.heads + head];
        sdB[i] = 0.f;
    }
    __syncthreads();
    auto load_pair = [&](const float* a_base, long long a_rs, float a_scale, const float* c_base, long long c_rs) {
        for (int i = threadIdx.x; i < n * (DH / 4); i += blockDim.x) {
            const int t = i / (DH / 4), c = (i - t * (DH / 4)) * 4;
            float4 a4 = *reinterpret_cast<const float4*>(a_base + (long long)sTok[t] * a_rs + head * DH + c);
            a4.x *= a_scale; a4.y *= a_scale; a4.z *= a_scale; a4.w *= a_scale;
            *reinterpret_cast<float4*>(sA + t * RS + c) = a4;
            *reinterpret_cast<float4*>(sC + t * RS + c) =
                *reinterpret_cast<const float4*>(c_base + (long long)sTok[t] * c_rs + head * DH + c);
        }
    };
    auto dot = [&](const float* row, const float (&v)[DH]) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(row + c);
            acc = fmaf(t.x, v[c], acc);
            acc = fmaf(t.y, v[c + 1], acc);
            acc = fmaf(t.z, v[c + 2], acc);
            acc = fmaf(t.w, v[c + 3], acc);
        }
        return acc;
    };
    auto load_row = [&](const float* g, float scale, float (&v)[DH]) {
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 t = *reinterpret_cast<const float4*>(g + c);
            v[c] = t.x * scale; v[c + 1] = t.y * scale; v[c + 2] = t.z * scale; v[c + 3] = t.w * scale;
        }
    };
    auto store_row = [&](long long off, const float (&v)[DH], float scale) {
#pragma unroll
        for (int c = 0; c < DH; c += 4)
            store_split4(p.dqkv, off + c, make_float4(v[c] * scale, v[c + 1] * scale, v[c + 2] * scale, v[c + 3] * scale));
    };
    const int s2 = 2 * p.w - 1;
    auto valid_key = [&](int lj) { return p.key_mask == nullptr || p.key_mask[b * p.L + lj] != 0; };
    // relative-position table entry of (query i, key j)
    auto bias_idx = [&](int i, int j) {
        const int li = i / ww, ri = i - li * ww, i1 = ri / p.w, i2 = ri - i1 * p.w;
        const int lj = j / ww, rj = j - lj * ww, k1 = rj / p.w, k2 = rj - k1 * p.w;
        return ((li - lj + p.L - 1) * s2 + (i1 - k1 + p.w - 1)) * s2 + (i2 - k2 + p.w - 1);
    };
    const int me = threadIdx.x;
    // ---------------------------------------------------------------- phase A: thread = query i
    load_pair(p.qkv + D, 3LL * D, 1.f, p.qkv + 2 * D, 3LL * D);   // K, V
    __syncthreads();
    if (me < n) {
        const int i = me;
        float q[DH], go[DH];
        load_row(p.qkv + (long long)sTok[i] * (3 * D) + head * DH, p.scale, q);
        load_row(p.dout + (long long)sTok[i] * D + head * DH, 1.f, go);
        float m = -INFINITY;
        for (int j = 0; j < n; ++j)
            if (valid_key(j / ww)) m = fmaxf(m, dot(sA + j * RS, q) + sB[bias_idx(i, j)]);
        float sum = 0.f, Dv = 0.f;
        for (int j = 0; j < n; ++j) {
            if (!valid_key(j / ww)) continue;
            const float e = __expf(dot(sA + j * RS, q) + sB[bias_idx(i, j)] - m);
            sum += e;
            Dv = fmaf(e, dot(sC + j * RS, go), Dv);
        }
        const float inv = 1.f / sum;
        Dv *= inv;
        sM[i] = m;
        sL[i] = inv;
        sD[i] = Dv;
        float dq[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) dq[c] = 0.f;
        for (int j = 0; j < n; ++j) {
            if (!valid_key(j / ww)) continue;
            const float pij = __expf(dot(sA + j * RS, q) + sB[bias_idx(i, j)] - m) * inv;
            const float ds = pij * (dot(sC + j * RS, go) - Dv);
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(sA + j * RS + c);
                dq[c] = fmaf(ds, t.x, dq[c]);
                dq[c + 1] = fmaf(ds, t.y, dq[c + 1]);
                dq[c + 2] = fmaf(ds, t.z, dq[c + 2]);
                dq[c + 3] = fmaf(ds, t.w, dq[c + 3]);
            }
        }
        store_row((long long)sTok[i] * (3 * D) + head * DH, dq, p.scale);
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase B: thread = key j
    load_pair(p.qkv, 3LL * D, p.scale, p.dout, (long long)D);     // q * scale, dO
    __syncthreads();
    if (me < n) {
        const int j = me;
        const bool ok = valid_key(j / ww);
        float k[DH], acc[DH];
        load_row(p.qkv + (long long)sTok[j] * (3 * D) + D + head * DH, 1.f, k);
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = 0.f;
        if (ok) {
            for (int i = 0; i < n; ++i) {                              // dV_j = sum_i P_ij dO_i
                const float pij = __expf(dot(sA + i * RS, k) + sB[bias_idx(i, j)] - sM[i]) * sL[i];
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(sC + i * RS + c);
                    acc[c] = fmaf(pij, t.x, acc[c]);
                    acc[c + 1] = fmaf(pij, t.y, acc[c + 1]);
                    acc[c + 2] = fmaf(pij, t.z, acc[c + 2]);
                    acc[c + 3] = fmaf(pij, t.w, acc[c + 3]);
                }
            }
        }
        store_row((long long)sTok[j] * (3 * D) + 2 * D + head * DH, acc, 1.f);
        float v[DH];
        load_row(p.qkv + (long long)sTok[j] * (3 * D) + 2 * D + head * DH, 1.f, v);
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = 0.f;
        if (ok) {
            for (int i = 0; i < n; ++i) {                              // dK_j = sum_i dS_ij (scale Q_i), bias gradient
                const int bi = bias_idx(i, j);
                const float pij = __expf(dot(sA + i * RS, k) + sB[bi] - sM[i]) * sL[i];
                const float ds = pij * (dot(sC + i * RS, v) - sD[i]);
                atomicAdd(&sdB[bi], ds);
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(sA + i * RS + c);
                    acc[c] = fmaf(ds, t.x, acc[c]);
                    acc[c + 1] = fmaf(ds, t.y, acc[c + 1]);
                    acc[c + 2] = fmaf(ds, t.z, acc[c + 2]);
                    acc[c + 3] = fmaf(ds, t.w, acc[c + 3]);
                }
            }
        }
        store_row((long long)sTok[j] * (3 * D) + D + head * DH, acc, 1.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (sdB[i] != 0.f) atomicAdd(&p.dbias[i * p.heads + head], sdB[i]);
}

}  // namespace a2x

extern "C" {

int a2x_layernorm_bwd(const float* x, int x_cs, const float* dy, int dy_cs, long long rows, int C, const float* gamma,
                      float eps, float* dx_accum, int dx_cs, double* dgamma, double* dbeta, a2x_stream_t stream) {
    A2X_REQUIRE(x && dy && gamma && dx_accum && dgamma && dbeta && rows > 0, "layernorm_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    long long b = (rows + 63) / 64;
    if (b > 148 * 4) b = 148 * 4;
    const int g = (int)(b < 1 ? 1 : b);
    if (C == 128) a2x::layernorm_bwd_kernel<1><<<g, 256, 0, st>>>(x, x_cs, dy, dy_cs, gamma, eps, dx_accum, dx_cs, rows, dgamma, dbeta);
    else if (C == 256) a2x::layernorm_bwd_kernel<2><<<g, 256, 0, st>>>(x, x_cs, dy, dy_cs, gamma, eps, dx_accum, dx_cs, rows, dgamma, dbeta);
    else {
        a2x::set_error("layernorm_bwd: channel count %d not in {128, 256}", C);
        return 1;
    }
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_gelu_fwd(const float* x, long long n, const a2x_output* y, a2x_stream_t stream) {
    A2X_REQUIRE(x && y && y->hi && n > 0 && n % 4 == 0, "gelu_fwd: bad args");
    long long b = (n / 4 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    a2x::gelu_fwd_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(x, n / 4, a2x::tr_split(y));
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_gelu_bwd(const float* dy, const float* x, long long n, const a2x_output* dx, a2x_stream_t stream) {
    A2X_REQUIRE(dy && x && dx && dx->hi && n > 0 && n % 4 == 0, "gelu_bwd: bad args");
    long long b = (n / 4 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    a2x::gelu_bwd_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(dy, x, n / 4, a2x::tr_split(dx));
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int a2x_window_attention_bwd(const float* qkv, const float* dout, const float* bias_table, const int* key_mask, int B,
                             int L, int H, int W, int heads, int dim_head, int window, int grid_mode, float scale,
                             float* dqkv, float* dbias_table, a2x_stream_t stream) {
    A2X_REQUIRE(dqkv, "window_attention_bwd: bad args");
    a2x_output o;
    o.hi = dqkv; o.b16 = nullptr; o.b16_plane = 0; o.cs = 3 * heads * dim_head;
    return a2x_window_attention_bwd_split(qkv, dout, bias_table, key_mask, B, L, H, W, heads, dim_head, window, grid_mode, scale,
                                          &o, dbias_table, stream);
}

int a2x_window_attention_bwd_split(const float* qkv, const float* dout, const float* bias_table, const int* key_mask, int B,
                                   int L, int H, int W, int heads, int dim_head, int window, int grid_mode, float scale,
                                   const a2x_output* dqkv, float* dbias_table, a2x_stream_t stream) {
    A2X_REQUIRE(qkv && dout && bias_table && dqkv && (dqkv->hi || dqkv->b16) && dbias_table && B > 0 && L > 0 && heads > 0 &&
                    window > 0,
                "window_attention_bwd: bad args");
    A2X_REQUIRE(dqkv->cs == 3 * heads * dim_head, "window_attention_bwd: dense [.., 3*heads*dim_head] gradient expected");
    A2X_REQUIRE(H % window == 0 && W % window == 0, "window_attention_bwd: H, W must be multiples of the window");
    a2x::WinAttBwdParams p;
    p.qkv = qkv; p.dout = dout; p.bias = bias_table; p.key_mask = key_mask; p.dbias = dbias_table;
    p.dqkv.hi = dqkv->hi; p.dqkv.b16 = (__nv_bfloat16*)dqkv->b16; p.dqkv.ps = dqkv->b16_plane;
    p.B = B; p.L = L; p.H = H; p.W = W; p.heads = heads; p.w = window; p.grid_mode = grid_mode; p.scale = scale;
    const int n = L * window * window;
    const int nb = (2 * L - 1) * (2 * window - 1) * (2 * window - 1);
    cudaStream_t st0 = (cudaStream_t)stream;
    if (n == 4 || n == 16) {  // small windows: 128 / n windows per CTA
        const int num_windows = B * (H / window) * (W / window);
        const int G = 128 / n;
        const size_t sms = (size_t)(4 * 128 * (dim_head + 4) + 2 * 128 * (n + 1) + nb) * sizeof(float);
        const long long gs = (long long)((num_windows + G - 1) / G) * heads;
#define A2X_WASB(DH, NT)                                                                                        \
    do {                                                                                                        \
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::window_attention_small_bwd_kernel<DH, NT>,                     \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sms));            \
        a2x::window_attention_small_bwd_kernel<DH, NT><<<(unsigned)gs, 128, sms, st0>>>(p, num_windows);        \
    } while (0)
        if (dim_head == 16 && n == 4) A2X_WASB(16, 4);
        else if (dim_head == 32 && n == 4) A2X_WASB(32, 4);
        else if (dim_head == 64 && n == 4) A2X_WASB(64, 4);
        else if (dim_head == 16) A2X_WASB(16, 16);
        else if (dim_head == 32) A2X_WASB(32, 16);
        else if (dim_head == 64) A2X_WASB(64, 16);
        else {
            a2x::set_error("window_attention_bwd: dim_head %d not in {16, 32, 64}", dim_head);
            return 1;
        }
#undef A2X_WASB
        A2X_LAUNCHED();
        A2X_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (a2x::wt_eligible(window, dim_head, n, L)) {
        a2x::WinTcParams q;
        q.qkv = qkv; q.dout = dout; q.bias = bias_table; q.key_mask = key_mask; q.out = p.dqkv; q.dbias = dbias_table;
        q.B = B; q.L = L; q.H = H; q.W = W; q.heads = heads; q.grid_mode = grid_mode; q.scale = scale;
        const int num_windows = B * (H / window) * (W / window);
        const int smem_tc = a2x::wt_smem_bytes(n, true);
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::window_attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc));
        a2x::window_attention_tc_kernel<true><<<a2x::wt_grid(num_windows, heads, 1), a2x::WT_THREADS, smem_tc, st0>>>(q, num_windows);
        A2X_LAUNCHED();
        A2X_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = (size_t)(4 * n * (dim_head + 4) + n * (n + 1) + 2 * nb + n) * sizeof(float);
    const long long grid = (long long)B * (H / window) * (W / window) * heads;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 200 * 1024) {  // the n x n matrix does not fit: recompute-from-rows kernel (n <= 256)
        const size_t sml = (size_t)(2 * n * (dim_head + 4) + 3 * n + 2 * nb + n) * sizeof(float);
        A2X_REQUIRE(n <= 256 && sml <= 200 * 1024, "window_attention_bwd: window of %d tokens does not fit shared memory", n);
#define A2X_WAL(DH)                                                                                                   \
    do {                                                                                                              \
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::window_attention_bwd_large_kernel<DH>,                               \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml));                  \
        a2x::window_attention_bwd_large_kernel<DH><<<(unsigned)grid, 256, sml, st>>>(p);                              \
    } while (0)
        if (dim_head == 16) A2X_WAL(16);
        else if (dim_head == 32) A2X_WAL(32);
        else if (dim_head == 64) A2X_WAL(64);
        else {
            a2x::set_error("window_attention_bwd: dim_head %d not in {16, 32, 64}", dim_head);
            return 1;
        }
#undef A2X_WAL
        A2X_LAUNCHED();
        A2X_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    if (dim_head == 32) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::window_attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a2x::window_attention_bwd_kernel<32><<<(unsigned)grid, 128, smem, st>>>(p);
    } else if (dim_head == 16) {
        A2X_CHECK_CUDA(cudaFuncSetAttribute(a2x::window_attention_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a2x::window_attention_bwd_kernel<16><<<(unsigned)grid, 128, smem, st>>>(p);
    } else {
        a2x::set_error("window_attention_bwd: dim_head %d not in {16, 32}", dim_head);
        return 1;
    }
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

namespace a2x {
// out[b][p][:] = mean_l x[b][l][p][:]   and its adjoint   dst[b][l][p][:] = scale * src[b][p][:]
__global__ void agent_mean_kernel(const float* __restrict__ x, int L, long long img4, float* __restrict__ out,
                                  long long total4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / img4, e = i - b * img4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) {
            const float4 t = reinterpret_cast<const float4*>(x)[(b * L + l) * img4 + e];
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        const float inv = 1.f / (float)L;
        reinterpret_cast<float4*>(out)[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
    }
}
__global__ void agent_broadcast_kernel(const float* __restrict__ src, int L, long long img4, float scale,
                                       float* __restrict__ dst, long long total4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long slot = i / img4, e = i - slot * img4;
        const float4 t = reinterpret_cast<const float4*>(src)[(slot / L) * img4 + e];
        reinterpret_cast<float4*>(dst)[i] = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
    }
}
}  // namespace a2x

extern "C" {
int a2x_agent_mean(const float* x, int B, int L, long long img_elems, float* out, a2x_stream_t stream) {
    A2X_REQUIRE(x && out && B > 0 && L > 0 && img_elems > 0 && img_elems % 4 == 0, "agent_mean: bad args");
    const long long total4 = (long long)B * (img_elems / 4);
    long long b = (total4 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    a2x::agent_mean_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(x, L, img_elems / 4, out, total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int a2x_agent_broadcast(const float* src, int B, int L, long long img_elems, float scale, float* dst, a2x_stream_t stream) {
    A2X_REQUIRE(src && dst && B > 0 && L > 0 && img_elems > 0 && img_elems % 4 == 0, "agent_broadcast: bad args");
    const long long total4 = (long long)B * L * (img_elems / 4);
    long long b = (total4 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    a2x::agent_broadcast_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(src, L, img_elems / 4, scale, dst, total4);
    A2X_LAUNCHED();
    A2X_CHECK_CUDA(cudaGetLastError());
    return 0;
}
}  // extern "C"
