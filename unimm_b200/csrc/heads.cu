// Heads, score assembly and losses: pooler + NSP head, log-sum-exp tails of the LM head,
// val_lm-style per-sequence scores, likelihood / unlikelihood, NSP and image-KL losses, and the
// boundary check that the caller's dense masks really are what the descriptors regenerate.
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
    if (warp == 0) t = warp_sum(t);
    if (threadIdx.x == 0) scratch[0] = t;
    __syncthreads();
    return scratch[0];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? scratch[threadIdx.x] : -INFINITY;
    if (warp == 0) t = warp_max(t);
    if (threadIdx.x == 0) scratch[0] = t;
    __syncthreads();
    return scratch[0];
}

// NSP head (reference models/vilbert_dialog.py:1062-1070, fusion 'mul'): one warp per sequence over the pooled vectors the two
// pooler GEMMs produced (engine.cu: pooler_head)
__global__ void __launch_bounds__(128)
nsp_from_pooled_kernel(const float* __restrict__ pt, const float* __restrict__ pv, int n, int Hb, const float* __restrict__ Wn,
                       const float* __restrict__ bn, float* __restrict__ nsp) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= n) return;
    const float4* t = reinterpret_cast<const float4*>(pt + static_cast<size_t>(b) * Hb);
    const float4* v = reinterpret_cast<const float4*>(pv + static_cast<size_t>(b) * Hb);
    const float4* w0 = reinterpret_cast<const float4*>(Wn);
    const float4* w1 = reinterpret_cast<const float4*>(Wn + Hb);
    float a0 = 0.f, a1 = 0.f;
    for (int i = lane; i < Hb / 4; i += 32) {
        const float4 x = t[i], y = v[i], p = __ldg(w0 + i), q = __ldg(w1 + i);
        const float zx = x.x * y.x, zy = x.y * y.y, zz = x.z * y.z, zw = x.w * y.w;
        a0 = fmaf(zx, p.x, a0); a0 = fmaf(zy, p.y, a0); a0 = fmaf(zz, p.z, a0); a0 = fmaf(zw, p.w, a0);
        a1 = fmaf(zx, q.x, a1); a1 = fmaf(zy, q.y, a1); a1 = fmaf(zz, q.z, a1); a1 = fmaf(zw, q.w, a1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) { nsp[b * 2] = a0 + bn[0]; nsp[b * 2 + 1] = a1 + bn[1]; }
}

// ---- LM-head backward helpers -----------------------------------------------------------------------------------------
// 32 x 32 tiles through shared memory: coalesced 16-bit reads along src rows, coalesced writes along dst rows
__global__ void __launch_bounds__(256)
transpose16_kernel(const uint16_t* __restrict__ src, int lds, int rows, int cols, uint16_t* __restrict__ dst, int ldt) {
    __shared__ uint16_t tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[static_cast<size_t>(r) * lds + c] : uint16_t(0);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) dst[static_cast<size_t>(c) * ldt + r] = tile[tx][i];
    }
}

// d loss_i / d z_ij = coef_i * (p_ij - [j == y_i])  (reference models/vilbert_dialog.py:1577-1595):
//   likelihood row (w > 0):      loss_i = -w log p_y                         coef = w
//   unlikelihood row (w == -1):  loss_i = -log(max(1 - p_y, 1e-6))           coef = -p_y / (1 - p_y), 0 where the clamp is active
__global__ void lm_loss_coef_kernel(const float* __restrict__ logp, const float* __restrict__ weight, int n, float scale, float* __restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w = weight[i];
    float c = 0.f;
    if (w > 0.f) c = w;
    else if (w == -1.f) {
        const float p = expf(logp[i]);
        c = (1.0f - p > 1e-6f) ? -p / (1.0f - p) : 0.f;
    }
    coef[i] = c * scale;
}

template <bool FP16>
__global__ void row_sums16_kernel(const uint16_t* __restrict__ x, int ld, int rows, int cols, float alpha, float* __restrict__ out) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) {
        const uint16_t v = x[static_cast<size_t>(r) * ld + c];
        s += FP16 ? __half2float(*reinterpret_cast<const __half*>(&v)) : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&v));
    }
    s = warp_sum(s);
    if (lane == 0) out[r] = s * alpha;
}

// ---- dgrad / wgrad helpers ------------------------------------------------------------------------------------------------
__global__ void amax_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ bits) {
    unsigned mx = 0u;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const unsigned u = __float_as_uint(fabsf(x[i]));
        mx = u > mx ? u : mx;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(bits, mx);
}
__global__ void scale_from_amax_kernel(const unsigned* __restrict__ bits, int want_scale, float* __restrict__ out2, float headroom) {
    const float amax = __uint_as_float(*bits) * headroom;
    float s = 1.f;
    if (want_scale && amax > 0.f && isfinite(amax)) {
        int e;
        frexpf(amax, &e);                    // amax = m * 2^e, m in [0.5, 1)
        s = ldexpf(1.f, 10 - e);             // amax * s in [2^9, 2^10): far from fp16's 65504 and from its subnormals
    }
    out2[0] = s;
    out2[1] = 1.f / s;
}
__device__ __forceinline__ float gelu_grad_fast(float v) {
    const float v2 = v * v;
    float r = fmaf(v2, kGeluC4, kGeluC3);
    r = fmaf(r, v2, kGeluC2);
    r = fmaf(r, v2, kGeluC1);
    r = fmaf(r, v2, kGeluC0);
    const float cdf = fast_rcp(1.0f + fast_ex2(r * v));
    return fmaf(v * 0.39894228040143267794f, fast_ex2(-0.72134752044448170368f * v2), cdf);
}

// One pass over a gradient matrix that is about to become a tensor-core operand: y = lp(v * scale) and colsum += v (the bias gradient),
// where v = x, or v = x * gelu'(t) when the matrix still has to go back through the erf GELU (models/vilbert_dialog.py:115-121) —
// the GELU backward, the cast and the bias gradient then read the fp32 gradient once and write 2 bytes per element instead of
// three reads and an fp32 write.
// GELU: 0 = none, 1 = pre-activation t in fp32, 2 = t as 16-bit values (encoding t_kind; the training forward's pre_act_lp epilogue)
template <int GELU>
__global__ void __launch_bounds__(256)
cast_colsum_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ t, int ldt, int rows, int cols, const float* __restrict__ scale,
                   bf16* __restrict__ y, int ldy, int lp_kind, float* __restrict__ colsum, DropArgs drop, int t_kind) {
    // CTA = 64 columns x 128 rows: a warp reads 256 contiguous bytes of one row (two columns per lane), the 8 warps interleave the rows
    __shared__ float part[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + lane * 2;
    const float s = scale[0];
    const int r0 = blockIdx.y * 128, r1 = min(rows, r0 + 128);
    float a0 = 0.f, a1 = 0.f;
    if (c < cols) {
#pragma unroll 4
        for (int r = r0 + warp; r < r1; r += 8) {
            float2 v = *reinterpret_cast<const float2*>(x + static_cast<size_t>(r) * ldx + c);
            if (drop.thresh != 0u) {        // the projection's output went through dropout: the same keep-mask on its gradient
                const uint32_t i0 = static_cast<uint32_t>(r) * static_cast<uint32_t>(cols) + static_cast<uint32_t>(c);
                v.x = drop_keep(drop.seed, i0, drop.thresh) ? v.x * drop.scale : 0.f;
                v.y = drop_keep(drop.seed, i0 + 1u, drop.thresh) ? v.y * drop.scale : 0.f;
            }
            if (GELU != 0) {
                // d/dt [t Phi(t)] = Phi(t) + t phi(t); Phi from the forward's rational fit (common.cuh gelu_fast: 3.5e-6), phi by one ex2
                float2 tv;
                if (GELU == 1) {
                    tv = *reinterpret_cast<const float2*>(t + static_cast<size_t>(r) * ldt + c);
                } else {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const bf16*>(t) + static_cast<size_t>(r) * ldt + c);
                    tv = t_kind == LP_FP16 ? __half22float2(*reinterpret_cast<const __half2*>(&w))
                                           : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
                }
                v.x *= gelu_grad_fast(tv.x);
                v.y *= gelu_grad_fast(tv.y);
            }
            a0 += v.x;
            a1 += v.y;
            *reinterpret_cast<uint32_t*>(y + static_cast<size_t>(r) * ldy + c) = pack_lp2(v.x * s, v.y * s, lp_kind);
        }
    }
    if (colsum == nullptr) return;
    part[warp][lane * 2] = a0;
    part[warp][lane * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < cols) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[w][threadIdx.x];
        atomicAdd(colsum + blockIdx.x * 64 + threadIdx.x, sum);
    }
}

__global__ void cast_scaled_kernel(const float* __restrict__ x, int ldx, int rows, int cols, const float* __restrict__ scale,
                                   bf16* __restrict__ y, int ldy, int lp_kind) {
    const float s = scale[0];
    const int cv = cols / 2;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < static_cast<size_t>(rows) * cv;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / cv;
        const int c = static_cast<int>(i % cv) * 2;
        const float2 v = *reinterpret_cast<const float2*>(x + r * ldx + c);
        *reinterpret_cast<uint32_t*>(y + r * ldy + c) = pack_lp2(v.x * s, v.y * s, lp_kind);
    }
}
// column sums of an [rows, cols] fp32 matrix (bias gradients): each CTA sums a 256-row slab of 128 columns (Kahan inside the slab)
// and adds its partial to the zeroed output with one atomic per column
__global__ void column_sums_kernel(const float* __restrict__ x, int ldx, int rows, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int r0 = blockIdx.y * 256, r1 = min(rows, r0 + 256);
    float s = 0.f, comp = 0.f;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
        const float v = x[static_cast<size_t>(r) * ldx + c] - comp;
        const float t = s + v;
        comp = (t - s) - v;
        s = t;
    }
    atomicAdd(out + c, s);
}

__global__ void segment_sum_kernel(const float* __restrict__ vals, const int* __restrict__ off, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int i = off[c]; i < off[c + 1]; ++i) s += vals[i];
    out[c] = s;
}

// log p(label) and log(max(1 - p(label), 1e-6)) from one row of materialised fp32 logits
__global__ void __launch_bounds__(256)
lse_logits_kernel(const float* __restrict__ logits, int ld, int V, const int* __restrict__ labels, float* __restrict__ logp,
                  float* __restrict__ ul) {
    __shared__ float scratch[32];
    const int row = blockIdx.x;
    const float* x = logits + static_cast<size_t>(row) * ld;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, x[i]);
    mx = block_max(mx, scratch);
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(x[i] - mx);
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        const int lab = labels[row];
        const float lp = (lab >= 0 && lab < V) ? (x[lab] - mx) - logf(s) : 0.f;
        logp[row] = lp;
        ul[row] = logf(fmaxf(1.0f - expf(lp), 1e-6f));  // reference :1587, clamp_min = 1e-6
    }
}

__global__ void lse_partials_kernel(const float2* __restrict__ partials, int tiles, const float* __restrict__ label_logit,
                                    int rows, float* __restrict__ logp, float* __restrict__ ul) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float2* p = partials + static_cast<size_t>(row) * tiles;
    float mx = -INFINITY;
    for (int i = lane; i < tiles; i += 32) mx = fmaxf(mx, p[i].x);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < tiles; i += 32) s += p[i].y * expf(p[i].x - mx);
    s = warp_sum(s);
    if (lane == 0) {
        const float lp = (label_logit[row] - mx) - logf(s);
        logp[row] = lp;
        ul[row] = logf(fmaxf(1.0f - expf(lp), 1e-6f));
    }
}

// Shared labelled rows (several (row, label) entries per LM-head row): log-sum-exp per unique row, then per entry the label's
// logit as a gathered dot product  h[urow] . E[label] + b[label]  (same 16-bit operands and fp32 accumulation as the fused GEMM
// epilogue, only the summation order differs) -> log p and log(max(1 - p, 1e-6)) (models/vilbert_dialog.py:1586-1591).
__global__ void lse_merge_kernel(const float2* __restrict__ partials, int tiles, int rows, float* __restrict__ lse) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float2* p = partials + static_cast<size_t>(row) * tiles;
    float mx = -INFINITY;
    for (int i = lane; i < tiles; i += 32) mx = fmaxf(mx, p[i].x);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < tiles; i += 32) s += p[i].y * expf(p[i].x - mx);
    s = warp_sum(s);
    if (lane == 0) lse[row] = mx + logf(s);
}

template <bool FP16>
__global__ void label_score_kernel(const bf16* __restrict__ h, int ldh, const bf16* __restrict__ E, int lde, const float* __restrict__ bias,
                                   const int* __restrict__ uidx, const int* __restrict__ labels, const float* __restrict__ lse, int n, int K,
                                   float* __restrict__ logp, float* __restrict__ ul) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int u = uidx[i], lab = labels[i];
    const uint4* a = reinterpret_cast<const uint4*>(h + static_cast<size_t>(u) * ldh);
    const uint4* b = reinterpret_cast<const uint4*>(E + static_cast<size_t>(lab) * lde);
    float acc = 0.f;
    for (int c = lane; c < K / 8; c += 32) {
        const uint4 x = a[c], y = __ldg(b + c);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 fx, fy;
            if (FP16) {
                fx = __half22float2(*reinterpret_cast<const __half2*>(&xs[j])); fy = __half22float2(*reinterpret_cast<const __half2*>(&ys[j]));
            } else {
                fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs[j])); fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ys[j]));
            }
            acc = fmaf(fx.x, fy.x, acc);
            acc = fmaf(fx.y, fy.y, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float lp = (acc + bias[lab]) - lse[u];
        logp[i] = lp;
        ul[i] = logf(fmaxf(1.0f - expf(lp), 1e-6f));
    }
}

// fp32 operands (fp32-class mode): the label logit as an fp32 dot product of the hidden row and the fp32 embedding row
__global__ void label_score_f32_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ E, int lde, const float* __restrict__ bias,
                                       const int* __restrict__ uidx, const int* __restrict__ labels, const float* __restrict__ lse, int n, int K,
                                       float* __restrict__ logp, float* __restrict__ ul) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int u = uidx[i], lab = labels[i];
    const float4* a = reinterpret_cast<const float4*>(h + static_cast<size_t>(u) * ldh);
    const float4* b = reinterpret_cast<const float4*>(E + static_cast<size_t>(lab) * lde);
    float acc = 0.f;
    for (int c = lane; c < K / 4; c += 32) {
        const float4 x = a[c], y = __ldg(b + c);
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float lp = (acc + bias[lab]) - lse[u];
        logp[i] = lp;
        ul[i] = logf(fmaxf(1.0f - expf(lp), 1e-6f));
    }
}

__global__ void scatter_scores_kernel(const float* __restrict__ logp, const float* __restrict__ ul,
                                      const int* __restrict__ flat_rows, int n, float* token_logp, float* token_ul) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (token_logp) token_logp[flat_rows[i]] = logp[i];
    if (token_ul) token_ul[flat_rows[i]] = ul[i];
}
// deterministic per-sequence sum: flat_rows is sorted, so sequence b owns a contiguous run
__global__ void seq_score_kernel(const float* __restrict__ logp, const int* __restrict__ flat_rows, int n, int B, int S,
                                 float* __restrict__ seq_score) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // lower bound of b*S in flat_rows
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (flat_rows[mid] < b * S) lo = mid + 1; else hi = mid;
    }
    float s = 0.f;
    for (int i = lo; i < n && flat_rows[i] < (b + 1) * S; ++i) s += logp[i];
    seq_score[b] = s;
}

// (sum_l w*(-logp) + sum_ul (-ul)) / #(lm_weight != 0)      (reference :1577-1595)
__global__ void __launch_bounds__(256)
lm_ul_loss_kernel(const float* __restrict__ logp, const float* __restrict__ ul, const int* __restrict__ flat_rows, int n,
                  const int64_t* __restrict__ lm_weight, int BS, float* __restrict__ out) {
    __shared__ float scratch[32];
    float acc = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long w = lm_weight[flat_rows[i]];
        if (w > 0) acc += -logp[i] * static_cast<float>(w);
        else if (w == -1) acc += -ul[i];
    }
    for (int i = threadIdx.x; i < BS; i += blockDim.x) cnt += (lm_weight[i] != 0) ? 1.f : 0.f;
    acc = block_sum(acc, scratch);
    cnt = block_sum(cnt, scratch);
    if (threadIdx.x == 0) out[0] = acc / cnt;
}

__global__ void __launch_bounds__(256) lm_ce_loss_kernel(const float* __restrict__ logp, int n, float* __restrict__ out) {
    __shared__ float scratch[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += -logp[i];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[0] = acc / static_cast<float>(n);
}

// F.cross_entropy(weight=w/w[0], reduction='mean') = sum_i w[y_i] * nll_i / sum_i w[y_i]   (reference :1605-1621)
__global__ void __launch_bounds__(256)
nsp_ce_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B,
                   const float* __restrict__ nsp_weight, float* __restrict__ out) {
    __shared__ float scratch[32];
    float w0 = 1.f, w1 = 1.f;
    if (nsp_weight != nullptr) { w0 = 1.f; w1 = nsp_weight[1] / nsp_weight[0]; }
    float num = 0.f, den = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float a = logits[2 * i], b = logits[2 * i + 1];
        const float mx = fmaxf(a, b);
        const float lse = mx + logf(expf(a - mx) + expf(b - mx));
        const long long y = labels[i];
        const float w = (y == 0) ? w0 : w1;
        num += w * (lse - (y == 0 ? a : b));
        den += w;
    }
    num = block_sum(num, scratch);
    den = block_sum(den, scratch);
    if (threadIdx.x == 0) out[0] = num / den;
}

// per selected (image_label == 1) region: sum_c t_c (log t_c - log_softmax(x)_c); atomics into acc[0], count in acc[1]
__global__ void __launch_bounds__(256)
image_kl_kernel(const float* __restrict__ v_logits, int ld, const float* __restrict__ target,
                const int64_t* __restrict__ image_label, int C, float* __restrict__ acc) {
    __shared__ float scratch[32];
    const int row = blockIdx.x;
    if (image_label[row] != 1) return;  // block-uniform
    const float* x = v_logits + static_cast<size_t>(row) * ld;
    const float* t = target + static_cast<size_t>(row) * C;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < C; i += blockDim.x) mx = fmaxf(mx, x[i]);
    mx = block_max(mx, scratch);
    float s = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s += expf(x[i] - mx);
    s = block_sum(s, scratch);
    const float lse = mx + logf(s);
    float kl = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        const float ti = t[i];
        if (ti > 0.f) kl += ti * (logf(ti) - (x[i] - lse));  // kl_div(input=logp, target): xlogy(t,t) - t*input
    }
    kl = block_sum(kl, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(acc, kl);
        atomicAdd(acc + 1, 1.0f);
    }
}
__global__ void image_kl_finish_kernel(const float* acc, float* out) { out[0] = acc[0] / acc[1]; }

// dense-mask check: one CTA per (sequence, row); compares the caller's mask row with the descriptor's
__global__ void __launch_bounds__(256)
verify_masks_kernel(const SeqDesc* __restrict__ desc, int S, int R, const void* __restrict__ txt_mask, int elem_bytes,
                    int is_2d, const int64_t* __restrict__ co_mask, int* __restrict__ mismatch) {
    const int b = blockIdx.y, r = blockIdx.x;
    const SeqDesc d = desc[b];
    int bad = 0;
    if (r < S) {
        int lo, hi, self;
        text_row_interval(d, r, S, lo, hi, self);
        for (int c = threadIdx.x; c < S; c += blockDim.x) {
            const int want = ((c >= lo && c < hi) || c == self) ? 1 : 0;
            long long got;
            const size_t idx = is_2d ? static_cast<size_t>(b) * S + c : (static_cast<size_t>(b) * S + r) * S + c;
            if (elem_bytes == 1) got = static_cast<const uint8_t*>(txt_mask)[idx];
            else got = static_cast<const int64_t*>(txt_mask)[idx];
            bad |= (got != want);
        }
    } else if (co_mask != nullptr) {  // rows S .. S+R-1 check the co-attention mask rows
        const int rr = r - S;
        int lo, hi;
        if (d.mode == 1) { lo = 0; hi = min(d.L, S); } else { lo = 1; hi = min(d.ctx, S); }
        for (int c = threadIdx.x; c < S; c += blockDim.x) {
            const int want = (c >= lo && c < hi) ? 1 : 0;
            bad |= (co_mask[(static_cast<size_t>(b) * R + rr) * S + c] != want);
        }
    }
    if (bad) atomicExch(mismatch, 1);
}

}  // namespace

int nsp_from_pooled(const float* pooled_t, const float* pooled_v, int n, int Hb, const float* Wn, const float* bn, float* nsp_logits,
                    cudaStream_t stream) {
    if (n == 0) return 0;
    UNIMM_CHECK(Hb % 4 == 0, "nsp head: bi_hidden_size must be a multiple of 4");
    nsp_from_pooled_kernel<<<(n + 3) / 4, 128, 0, stream>>>(pooled_t, pooled_v, n, Hb, Wn, bn, nsp_logits);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int transpose_16(const bf16* src, int lds, int rows, int cols, bf16* dst, int ldt, cudaStream_t stream) {
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(src), lds, rows, cols, reinterpret_cast<uint16_t*>(dst), ldt);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int lm_loss_coef(const float* logp, const float* weight, int n, float scale, float* coef, cudaStream_t stream) {
    if (n == 0) return 0;
    lm_loss_coef_kernel<<<(n + 255) / 256, 256, 0, stream>>>(logp, weight, n, scale, coef);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int row_sums_16(const bf16* x, int ld, int rows, int cols, int lp_kind, float alpha, float* out, cudaStream_t stream) {
    if (rows == 0) return 0;
    if (lp_kind == LP_FP16) row_sums16_kernel<true><<<(rows + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const uint16_t*>(x), ld, rows, cols, alpha, out);
    else row_sums16_kernel<false><<<(rows + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const uint16_t*>(x), ld, rows, cols, alpha, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int amax_scale(const float* x, size_t n, int want_scale, float* out2, cudaStream_t stream, const float* known_amax, float headroom) {
    if (known_amax != nullptr) {      // the producer of x already left max |x| (as float bits) in device memory: no pass over x
        scale_from_amax_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<const unsigned*>(known_amax), want_scale, out2, headroom);
        UNIMM_LAUNCH_CHECK(1);
        return 0;
    }
    // out2[0] doubles as the atomicMax cell before it receives the scale
    UNIMM_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(float), stream));
    int grid = static_cast<int>((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    amax_kernel<<<grid < 1 ? 1 : grid, 256, 0, stream>>>(x, n, reinterpret_cast<unsigned*>(out2 + 1));
    scale_from_amax_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<const unsigned*>(out2 + 1), want_scale, out2, headroom);
    UNIMM_LAUNCH_CHECK(2);
    return 0;
}

int cast_colsum_lp(const float* x, int ldx, const float* gelu_t, int ldt, int rows, int cols, const float* scale, bf16* y, int ldy, int lp_kind,
                   float* colsum, cudaStream_t stream, DropArgs drop, int gelu_t_kind) {
    UNIMM_CHECK(rows > 0 && cols % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0 && (gelu_t == nullptr || ldt % 2 == 0), "cast_colsum: even columns and leading dimensions");
    if (colsum != nullptr) UNIMM_CUDA_CHECK(cudaMemsetAsync(colsum, 0, sizeof(float) * cols, stream));
    const dim3 grid((cols + 63) / 64, (rows + 127) / 128);
    if (gelu_t != nullptr && gelu_t_kind >= 0)
        cast_colsum_kernel<2><<<grid, 256, 0, stream>>>(x, ldx, gelu_t, ldt, rows, cols, scale, y, ldy, lp_kind, colsum, drop, gelu_t_kind);
    else if (gelu_t != nullptr) cast_colsum_kernel<1><<<grid, 256, 0, stream>>>(x, ldx, gelu_t, ldt, rows, cols, scale, y, ldy, lp_kind, colsum, drop, -1);
    else cast_colsum_kernel<0><<<grid, 256, 0, stream>>>(x, ldx, nullptr, 0, rows, cols, scale, y, ldy, lp_kind, colsum, drop, -1);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int cast_scaled_lp(const float* x, int ldx, int rows, int cols, const float* scale, bf16* y, int ldy, int lp_kind, cudaStream_t stream) {
    UNIMM_CHECK(rows > 0 && cols % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0, "cast_scaled: even columns and leading dimensions");
    const size_t n = static_cast<size_t>(rows) * (cols / 2);
    int grid = static_cast<int>((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    cast_scaled_kernel<<<grid, 256, 0, stream>>>(x, ldx, rows, cols, scale, y, ldy, lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int column_sums_f32(const float* x, int ldx, int rows, int cols, float* out, cudaStream_t stream) {
    UNIMM_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * cols, stream));
    column_sums_kernel<<<dim3((cols + 127) / 128, (rows + 255) / 256), 128, 0, stream>>>(x, ldx, rows, cols, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int segment_sum(const float* vals, const int* off, int C, float* out, cudaStream_t stream) {
    if (C == 0) return 0;
    segment_sum_kernel<<<(C + 127) / 128, 128, 0, stream>>>(vals, off, C, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int lse_from_logits(const float* logits, int ld, int rows, int V, const int* labels, float* logp, float* ul,
                    cudaStream_t stream) {
    if (rows == 0) return 0;
    lse_logits_kernel<<<rows, 256, 0, stream>>>(logits, ld, V, labels, logp, ul);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int lse_from_partials(const float2* partials, int tiles, const float* label_logit, int rows, float* logp, float* ul,
                      cudaStream_t stream) {
    if (rows == 0) return 0;
    lse_partials_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(partials, tiles, label_logit, rows, logp, ul);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int lse_merge(const float2* partials, int tiles, int rows, float* lse, cudaStream_t stream) {
    if (rows == 0) return 0;
    lse_merge_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(partials, tiles, rows, lse);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int label_scores(const bf16* h, int ldh, const bf16* E, int lde, const float* bias, const int* uidx, const int* labels, const float* lse,
                 int n, int K, int lp_kind, float* logp, float* ul, cudaStream_t stream) {
    if (n == 0) return 0;
    UNIMM_CHECK(K % 8 == 0 && ldh % 8 == 0 && lde % 8 == 0, "label_scores: rows must be 16-byte aligned");
    if (lp_kind == LP_FP16) label_score_kernel<true><<<(n + 3) / 4, 128, 0, stream>>>(h, ldh, E, lde, bias, uidx, labels, lse, n, K, logp, ul);
    else label_score_kernel<false><<<(n + 3) / 4, 128, 0, stream>>>(h, ldh, E, lde, bias, uidx, labels, lse, n, K, logp, ul);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int label_scores_f32(const float* h, int ldh, const float* E, int lde, const float* bias, const int* uidx, const int* labels, const float* lse,
                     int n, int K, float* logp, float* ul, cudaStream_t stream) {
    if (n == 0) return 0;
    UNIMM_CHECK(K % 4 == 0 && ldh % 4 == 0 && lde % 4 == 0, "label_scores: rows must be 16-byte aligned");
    label_score_f32_kernel<<<(n + 3) / 4, 128, 0, stream>>>(h, ldh, E, lde, bias, uidx, labels, lse, n, K, logp, ul);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int scatter_scores(const float* logp, const float* ul, const int* flat_rows, int n, int B, int S, float* token_logp,
                   float* token_ul, float* seq_score, cudaStream_t stream) {
    if (token_logp) UNIMM_CUDA_CHECK(cudaMemsetAsync(token_logp, 0, sizeof(float) * B * S, stream));
    if (token_ul) UNIMM_CUDA_CHECK(cudaMemsetAsync(token_ul, 0, sizeof(float) * B * S, stream));
    if (n > 0 && (token_logp || token_ul))
        scatter_scores_kernel<<<(n + 255) / 256, 256, 0, stream>>>(logp, ul, flat_rows, n, token_logp, token_ul);
    if (seq_score) seq_score_kernel<<<(B + 127) / 128, 128, 0, stream>>>(logp, flat_rows, n, B, S, seq_score);
    UNIMM_LAUNCH_CHECK(2);
    return 0;
}

int lm_ul_loss(const float* logp, const float* ul, const int* flat_rows, int n, const int64_t* lm_weight, int BS,
               float* out, cudaStream_t stream) {
    lm_ul_loss_kernel<<<1, 256, 0, stream>>>(logp, ul, flat_rows, n, lm_weight, BS, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int lm_ce_loss(const float* logp, int n, float* out, cudaStream_t stream) {
    lm_ce_loss_kernel<<<1, 256, 0, stream>>>(logp, n, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int nsp_ce_loss(const float* nsp_logits, const int64_t* labels, int B, const float* nsp_weight, float* out,
                cudaStream_t stream) {
    nsp_ce_loss_kernel<<<1, 256, 0, stream>>>(nsp_logits, labels, B, nsp_weight, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int image_kl_loss(const float* v_logits, int ld, const float* target, const int64_t* image_label, int rows, int C, float* out,
                  cudaStream_t stream) {
    // out[1..2] is scratch for (sum, count)
    UNIMM_CUDA_CHECK(cudaMemsetAsync(out + 1, 0, 2 * sizeof(float), stream));
    image_kl_kernel<<<rows, 256, 0, stream>>>(v_logits, ld, target, image_label, C, out + 1);
    image_kl_finish_kernel<<<1, 1, 0, stream>>>(out + 1, out);
    UNIMM_LAUNCH_CHECK(2);
    return 0;
}

int verify_masks(const SeqDesc* desc, int B, int S, int R, const void* txt_mask, int txt_elem_bytes, int txt_is_2d,
                 const int64_t* co_mask, int* mismatch_flag, cudaStream_t stream) {
    UNIMM_CHECK(txt_elem_bytes == 1 || txt_elem_bytes == 8, "attention_mask must be bool/uint8 or int64");
    UNIMM_CHECK(!txt_is_2d, "2-D attention masks are not produced by the reference encoders");
    dim3 grid(S + (co_mask ? R : 0), B);
    verify_masks_kernel<<<grid, 256, 0, stream>>>(desc, S, R, txt_mask, txt_elem_bytes, txt_is_2d, co_mask, mismatch_flag);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace unimm
