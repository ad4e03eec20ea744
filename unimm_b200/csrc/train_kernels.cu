// Small HBM-bound kernels of the training step (reference train.py:445-463): everything around the tensor-core
// GEMMs / attention that the backward and the optimizer need and that is not already in rowwise.cu / heads.cu — the embedding sum and
// its scatter-add backward, the forward GELU from a saved pre-activation, row gather / scatter-add, element-wise helpers, the
// gradients of the NSP cross entropy and the masked image KL, and AdamW (pytorch_transformers' variant, see unimm_t_adamw) — plus
// their C-ABI entry points (include/unimm_b200.h, "training step").
#include <cmath>

#include "../../include/unimm_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace unimm {

size_t attention_backward_scratch(int B, int heads, int D, int Sq);
int attention_backward_lp(const AttnArgs& f, const float* dO, int lddo, const float* lse, float* dq, int lddq, float* dk, int lddk,
                          float* dv, int lddv, void* scratch, size_t scratch_bytes, cudaStream_t stream, float* amax_accum, const float* dO_amax);

namespace {

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : 0.f;
    if (warp == 0) r = warp_sum(r);
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    return scratch[0];
}
__device__ __forceinline__ float block_max_256(float v, float* scratch) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : -INFINITY;
    if (warp == 0) r = warp_max(r);
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    return scratch[0];
}

// ---- text embeddings: word + position + (type | type-extension), no LayerNorm (its input is what the backward needs)
__global__ void __launch_bounds__(128)
embed_sum_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids, const int64_t* __restrict__ pos_ids, int rows, int H,
                 int vocab, int max_pos, int type_vocab, int type_ext, const float* __restrict__ word_emb, const float* __restrict__ pos_emb,
                 const float* __restrict__ type_emb, const float* __restrict__ type_ext_emb, float* __restrict__ out, int* err_flag) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    long long id = ids[row], ty = type_ids[row], pos = pos_ids[row];
    if (id < 0 || id >= vocab || pos < 0 || pos >= max_pos || ty < 0 || ty >= type_vocab + type_ext) {
        if (lane == 0 && err_flag != nullptr) atomicExch(err_flag, 1);
        id = min(max(id, 0LL), (long long)vocab - 1);
        pos = min(max(pos, 0LL), (long long)max_pos - 1);
        ty = min(max(ty, 0LL), (long long)(type_vocab + type_ext - 1));
    }
    const float* w = word_emb + static_cast<size_t>(id) * H;
    const float* p = pos_emb + static_cast<size_t>(pos) * H;
    const float* t = (ty < type_vocab) ? type_emb + static_cast<size_t>(ty) * H : type_ext_emb + static_cast<size_t>(ty - type_vocab) * H;
    for (int c = lane * 4; c < H; c += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + c));
        const float4 d = __ldg(reinterpret_cast<const float4*>(t + c));
        *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * H + c) =
            make_float4((a.x + b.x) + d.x, (a.y + b.y) + d.y, (a.z + b.z) + d.z, (a.w + b.w) + d.w);
    }
}

// backward: the row's gradient is added to the three table rows it was gathered from.  Rows whose gradient is exactly zero (padding
// and every position nothing labelled can see) are skipped: most of a batch.
__global__ void __launch_bounds__(128)
embed_bwd_kernel(const float* __restrict__ d, const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                 const int64_t* __restrict__ pos_ids, int rows, int H, int type_vocab, float* __restrict__ d_word, float* __restrict__ d_pos,
                 float* __restrict__ d_type, float* __restrict__ d_type_ext) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long id = ids[row], ty = type_ids[row], pos = pos_ids[row];
    float* w = d_word + static_cast<size_t>(id) * H;
    float* p = d_pos + static_cast<size_t>(pos) * H;
    float* t = (ty < type_vocab) ? d_type + static_cast<size_t>(ty) * H : d_type_ext + static_cast<size_t>(ty - type_vocab) * H;
    for (int c = lane * 4; c < H; c += 128) {
        const float4 g = *reinterpret_cast<const float4*>(d + static_cast<size_t>(row) * H + c);
        if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
        atomicAdd(w + c, g.x); atomicAdd(w + c + 1, g.y); atomicAdd(w + c + 2, g.z); atomicAdd(w + c + 3, g.w);
        atomicAdd(p + c, g.x); atomicAdd(p + c + 1, g.y); atomicAdd(p + c + 2, g.z); atomicAdd(p + c + 3, g.w);
        atomicAdd(t + c, g.x); atomicAdd(t + c + 1, g.y); atomicAdd(t + c + 2, g.z); atomicAdd(t + c + 3, g.w);
    }
}

// g = gelu_erf(t) as the 16-bit operand of the next GEMM (and / or fp32, the input of the heads' LayerNorm)
__global__ void gelu_fwd_lp_kernel(const float* __restrict__ t, size_t n4, float* __restrict__ g32, bf16* __restrict__ g, int lp_kind) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(t)[i];
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        if (g32 != nullptr) reinterpret_cast<float4*>(g32)[i] = v;
        if (g != nullptr) {
            uint2 p;
            p.x = pack_lp2(v.x, v.y, lp_kind);
            p.y = pack_lp2(v.z, v.w, lp_kind);
            reinterpret_cast<uint2*>(g)[i] = p;
        }
    }
}

// likelihood / unlikelihood loss value from the per-row log p (reference :1577-1595): out[0] = scale * (sum_{w > 0} -w log p +
// sum_{w == -1} -log(max(1 - p, 1e-6)))
__global__ void __launch_bounds__(256)
lm_ul_value_kernel(const float* __restrict__ logp, const float* __restrict__ w, int n, float scale, float* __restrict__ out) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float wi = w[i], lp = logp[i];
        if (wi > 0.f) s -= wi * lp;
        else if (wi == -1.f) s -= logf(fmaxf(1.0f - expf(lp), 1e-6f));
    }
    s = block_sum_256(s, scratch);
    if (threadIdx.x == 0) out[0] = s * scale;
}

// y = x o keep-mask / (1 - p) (forward and, applied to the gradient, backward of nn.Dropout); y32 may alias x
__global__ void dropout_kernel(const float* x, size_t n, DropArgs d, float* y32, bf16* __restrict__ y16, int lp_kind) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = drop_keep(d.seed, static_cast<uint32_t>(i), d.thresh) ? x[i] * d.scale : 0.f;
        if (y32 != nullptr) y32[i] = v;
        if (y16 != nullptr) y16[i] = lp_from_f32(v, lp_kind);
    }
}

// element-wise helpers on fp32 vectors: 0: out = a + b, 1: out = a * b, 2: out = a * [b > 0] (ReLU backward from the ReLU's output),
// 3: out = alpha * a, 4: out = a + alpha * b
__global__ void ew_kernel(int op, size_t n, const float* a, const float* b, float* out, float alpha) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float x = a[i];
        float r;
        if (op == 0) r = x + b[i];
        else if (op == 1) r = x * b[i];
        else if (op == 2) r = b[i] > 0.f ? x : 0.f;
        else if (op == 3) r = alpha * x;
        else r = x + alpha * b[i];
        out[i] = r;
    }
}

__global__ void gather_rows_f32_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int H, float* __restrict__ dst) {
    const float* s = src + static_cast<size_t>(idx[blockIdx.x]) * lds;
    for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + static_cast<size_t>(blockIdx.x) * H + c) = *reinterpret_cast<const float4*>(s + c);
}
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int H, float* __restrict__ dst, int ldd) {
    float* d = dst + static_cast<size_t>(idx[blockIdx.x]) * ldd;
    for (int c = threadIdx.x; c < H; c += blockDim.x) atomicAdd(d + c, src[static_cast<size_t>(blockIdx.x) * H + c]);
}

// weighted NSP cross entropy (reference :1605-1621): loss = sum_i w[y_i] (lse_i - x_i[y_i]) / sum_i w[y_i], w = nsp_weight / nsp_weight[0];
// d logits = grad_scale * w[y_i] (softmax_i - onehot_i) / sum w
__global__ void __launch_bounds__(256)
nsp_ce_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B, const float* __restrict__ nsp_weight,
                  float grad_scale, float* __restrict__ loss, float* __restrict__ dlogits) {
    __shared__ float scratch[32];
    const float w1 = nsp_weight != nullptr ? nsp_weight[1] / nsp_weight[0] : 1.f;
    float num = 0.f, den = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float a = logits[2 * i], b = logits[2 * i + 1];
        const float mx = fmaxf(a, b);
        const float lse = mx + logf(expf(a - mx) + expf(b - mx));
        const long long y = labels[i];
        const float w = (y == 0) ? 1.f : w1;
        num += w * (lse - (y == 0 ? a : b));
        den += w;
    }
    num = block_sum_256(num, scratch);
    den = block_sum_256(den, scratch);
    if (threadIdx.x == 0 && loss != nullptr) loss[0] = num / den;
    if (dlogits == nullptr) return;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float a = logits[2 * i], b = logits[2 * i + 1];
        const float mx = fmaxf(a, b);
        const float ea = expf(a - mx), eb = expf(b - mx), inv = 1.f / (ea + eb);
        const long long y = labels[i];
        const float f = grad_scale * ((y == 0) ? 1.f : w1) / den;
        dlogits[2 * i] = f * (ea * inv - (y == 0 ? 1.f : 0.f));
        dlogits[2 * i + 1] = f * (eb * inv - (y == 0 ? 0.f : 1.f));
    }
}

// masked image KL (reference :1569-1574): loss = sum_{rows with image_label == 1} sum_c t_c (log t_c - log_softmax(x)_c) / #selected;
// d x_c = grad_scale * (softmax(x)_c * sum_c' t_c' - t_c) / #selected on the selected rows, 0 elsewhere.  acc = (sum, count).
__global__ void __launch_bounds__(256)
image_kl_fwd_kernel(const float* __restrict__ x_all, int ld, const float* __restrict__ target, const int* __restrict__ target_row,
                    const int64_t* __restrict__ image_label, int C, float* __restrict__ acc) {
    __shared__ float scratch[32];
    const int row = blockIdx.x;
    if (image_label[row] != 1) return;
    const float* x = x_all + static_cast<size_t>(row) * ld;
    const float* t = target + static_cast<size_t>(target_row ? target_row[row] : row) * C;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < C; i += blockDim.x) mx = fmaxf(mx, x[i]);
    mx = block_max_256(mx, scratch);
    float s = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s += expf(x[i] - mx);
    s = block_sum_256(s, scratch);
    const float lse = mx + logf(s);
    float kl = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        const float ti = t[i];
        if (ti > 0.f) kl += ti * (logf(ti) - (x[i] - lse));
    }
    kl = block_sum_256(kl, scratch);
    if (threadIdx.x == 0) { atomicAdd(acc, kl); atomicAdd(acc + 1, 1.0f); }
}
__global__ void __launch_bounds__(256)
image_kl_bwd_kernel(const float* __restrict__ x_all, int ld, const float* __restrict__ target, const int* __restrict__ target_row,
                    const int64_t* __restrict__ image_label, int C, int ldd, const float* __restrict__ acc, float grad_scale,
                    float* __restrict__ loss, float* __restrict__ dx_all) {
    __shared__ float scratch[32];
    const int row = blockIdx.x;
    if (row == 0 && threadIdx.x == 0 && loss != nullptr) loss[0] = acc[0] / acc[1];
    float* dx = dx_all + static_cast<size_t>(row) * ldd;
    if (image_label[row] != 1) {
        for (int i = threadIdx.x; i < ldd; i += blockDim.x) dx[i] = 0.f;
        return;
    }
    const float* x = x_all + static_cast<size_t>(row) * ld;
    const float* t = target + static_cast<size_t>(target_row ? target_row[row] : row) * C;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < C; i += blockDim.x) mx = fmaxf(mx, x[i]);
    mx = block_max_256(mx, scratch);
    float s = 0.f, ts = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s += expf(x[i] - mx); ts += t[i]; }
    s = block_sum_256(s, scratch);
    ts = block_sum_256(ts, scratch);
    const float f = grad_scale / acc[1], inv = 1.f / s;
    for (int i = threadIdx.x; i < ldd; i += blockDim.x) dx[i] = i < C ? f * (expf(x[i] - mx) * inv * ts - t[i]) : 0.f;
}

// AdamW as pytorch_transformers.optimization.AdamW.step does it (the optimizer of train.py:347; third-party, not in /root/reference —
// restated from its published source, oracle/adamw.py is the test-side copy):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= step_size * m / (sqrt(v) + eps),  step_size = lr sqrt(1 - b2^t) / (1 - b1^t);
//   then the decoupled decay  p -= lr * wd * p  on the UPDATED p.
// inv_grad_scale multiplies g first (loss scaling / accumulation).  The 16-bit operand copy of the parameter is refreshed in the same pass.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n, float lr,
                             float b1, float b2, float eps, float wd, float step_size, float inv_grad_scale, bf16* __restrict__ p_lp, int lp_kind) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * inv_grad_scale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float pi = p[i] - step_size * mi / (sqrtf(vi) + eps);
        if (wd > 0.f) pi -= lr * wd * pi;
        p[i] = pi;
        if (p_lp != nullptr) p_lp[i] = lp_from_f32(pi, lp_kind);
    }
}

inline int grid_for(size_t n, int per_block = 256, int cap = 148 * 16) {
    size_t g = (n + per_block - 1) / per_block;
    if (g > static_cast<size_t>(cap)) g = cap;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace
}  // namespace unimm

using namespace unimm;

extern "C" {

int unimm_t_embed_text_sum(const int64_t* d_ids, const int64_t* d_type_ids, const int64_t* d_pos_ids, int rows, int H, int vocab, int max_pos,
                           int type_vocab, int type_ext, const float* d_word, const float* d_pos, const float* d_type, const float* d_type_ext,
                           float* d_out, int* d_err_flag, void* stream) {
    UNIMM_CHECK(d_ids && d_type_ids && d_pos_ids && d_word && d_pos && d_type && d_type_ext && d_out && rows > 0 && H % 4 == 0, "bad argument");
    embed_sum_kernel<<<(rows + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_ids, d_type_ids, d_pos_ids, rows, H, vocab, max_pos, type_vocab,
                                                                                    type_ext, d_word, d_pos, d_type, d_type_ext, d_out, d_err_flag);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_embed_text_backward(const float* d_dsum, const int64_t* d_ids, const int64_t* d_type_ids, const int64_t* d_pos_ids, int rows, int H,
                                int type_vocab, float* d_dword, float* d_dpos, float* d_dtype, float* d_dtype_ext, void* stream) {
    UNIMM_CHECK(d_dsum && d_ids && d_type_ids && d_pos_ids && d_dword && d_dpos && d_dtype && d_dtype_ext && rows > 0 && H % 4 == 0, "bad argument");
    embed_bwd_kernel<<<(rows + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_dsum, d_ids, d_type_ids, d_pos_ids, rows, H, type_vocab, d_dword,
                                                                                    d_dpos, d_dtype, d_dtype_ext);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_gelu(const float* d_t, int64_t n, float* d_g_f32, void* d_g_lp, int lp_kind, void* stream) {
    UNIMM_CHECK(d_t && (d_g_lp || d_g_f32) && n > 0 && (n & 3) == 0, "gelu: element count must be a positive multiple of 4");
    gelu_fwd_lp_kernel<<<grid_for(static_cast<size_t>(n) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_t, static_cast<size_t>(n) / 4, d_g_f32,
                                                                                                            static_cast<bf16*>(d_g_lp), lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_lm_ul_value(const float* d_logp, const float* d_weight, int n, float scale, float* d_out, void* stream) {
    UNIMM_CHECK(d_logp && d_weight && d_out && n > 0, "bad argument");
    lm_ul_value_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_logp, d_weight, n, scale, d_out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_dropout(const float* d_x, int64_t n, uint32_t seed, float p, float* d_y_f32, void* d_y_lp, int lp_kind, void* stream) {
    UNIMM_CHECK(d_x && (d_y_f32 || d_y_lp) && n > 0 && n < (int64_t(1) << 32) && p > 0.f && p < 1.f, "dropout: 0 < p < 1, fewer than 2^32 elements");
    dropout_kernel<<<grid_for(static_cast<size_t>(n)), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, static_cast<size_t>(n), make_drop(seed, p), d_y_f32,
                                                                                                    static_cast<bf16*>(d_y_lp), lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_gemm_drop(const void* d_A, int lda, const void* d_W, int ldw, int M, int N, int K, const float* d_bias, const float* d_residual, int ldr,
                      uint32_t drop_seed, float drop_p, float* d_out_f32, int ldo_f32, int lp_kind, void* stream) {
    GemmEpilogue ep;
    ep.lp_kind = lp_kind; ep.bias = d_bias; ep.residual = d_residual; ep.ldr = ldr; ep.out_f32 = d_out_f32; ep.ldo_f32 = ldo_f32;
    ep.drop = make_drop(drop_seed, drop_p);
    return gemm_umma_bf16(static_cast<const bf16*>(d_A), lda, static_cast<const bf16*>(d_W), ldw, M, N, K, ep, 0, 0, static_cast<cudaStream_t>(stream));
}

int unimm_t_ew(int op, int64_t n, const float* d_a, const float* d_b, float* d_out, float alpha, void* stream) {
    UNIMM_CHECK(op >= 0 && op <= 4 && n > 0 && d_a && d_out && (op == 3 || d_b), "bad argument");
    ew_kernel<<<grid_for(static_cast<size_t>(n)), 256, 0, static_cast<cudaStream_t>(stream)>>>(op, static_cast<size_t>(n), d_a, d_b, d_out, alpha);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_gather_rows(const float* d_src, int lds, const int32_t* d_idx, int n, int H, float* d_dst, void* stream) {
    UNIMM_CHECK(d_src && d_idx && d_dst && n > 0 && H % 4 == 0 && lds % 4 == 0, "bad argument");
    gather_rows_f32_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_src, lds, d_idx, H, d_dst);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_scatter_add_rows(const float* d_src, const int32_t* d_idx, int n, int H, float* d_dst, int ldd, void* stream) {
    UNIMM_CHECK(d_src && d_idx && d_dst && n > 0, "bad argument");
    scatter_add_rows_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_idx, H, d_dst, ldd);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_nsp_ce(const float* d_logits, const int64_t* d_labels, int B, const float* d_nsp_weight, float grad_scale, float* d_loss,
                   float* d_dlogits, void* stream) {
    UNIMM_CHECK(d_logits && d_labels && B > 0, "bad argument");
    nsp_ce_bwd_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_logits, d_labels, B, d_nsp_weight, grad_scale, d_loss, d_dlogits);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_image_kl(const float* d_logits, int ld, const float* d_target, const int32_t* d_target_row, const int64_t* d_image_label, int rows,
                     int C, float grad_scale, float* d_loss, float* d_dlogits, int ldd, float* d_acc2, void* stream) {
    UNIMM_CHECK(d_logits && d_target && d_image_label && d_acc2 && rows > 0 && C > 0 && ld >= C, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    UNIMM_CUDA_CHECK(cudaMemsetAsync(d_acc2, 0, 2 * sizeof(float), st));
    image_kl_fwd_kernel<<<rows, 256, 0, st>>>(d_logits, ld, d_target, d_target_row, d_image_label, C, d_acc2);
    UNIMM_LAUNCH_CHECK(1);
    if (d_dlogits != nullptr) {
        UNIMM_CHECK(ldd >= C, "image KL: gradient leading dimension smaller than the class count");
        image_kl_bwd_kernel<<<rows, 256, 0, st>>>(d_logits, ld, d_target, d_target_row, d_image_label, C, ldd, d_acc2, grad_scale, d_loss, d_dlogits);
        UNIMM_LAUNCH_CHECK(1);
    } else if (d_loss != nullptr) {
        UNIMM_CHECK(false, "image KL: the loss is written by the gradient pass; pass d_dlogits");
    }
    return 0;
}

int unimm_t_adamw(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, int correct_bias, float inv_grad_scale, void* d_p_lp, int lp_kind, void* stream) {
    UNIMM_CHECK(d_p && d_g && d_m && d_v && n > 0 && step >= 1, "bad argument");
    float step_size = lr;
    if (correct_bias) step_size = static_cast<float>(lr * std::sqrt(1.0 - std::pow(static_cast<double>(beta2), step)) /
                                                     (1.0 - std::pow(static_cast<double>(beta1), step)));
    adamw_kernel<<<grid_for(static_cast<size_t>(n)), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_p, d_g, d_m, d_v, static_cast<size_t>(n), lr, beta1,
                                                                                                 beta2, eps, weight_decay, step_size, inv_grad_scale,
                                                                                                 static_cast<bf16*>(d_p_lp), lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_k_attention_lse(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo, int B, int heads,
                          int D, int Sq, int Skv, int mask_kind, const unimm_seq_desc_t* d_desc, const float* d_key_mask, int lp_kind,
                          float* d_lse, uint32_t drop_seed, float drop_p, void* stream) {
    UNIMM_CHECK(lp_kind == LP_BF16 || lp_kind == LP_FP16, "attention: 16-bit tensors only");
    AttnArgs a;
    a.q = d_q; a.ldq = ldq; a.k = d_k; a.ldk = ldk; a.v = d_v; a.ldv = ldv; a.o = d_o; a.ldo = ldo;
    a.B = B; a.heads = heads; a.D = D; a.Sq = Sq; a.Skv = Skv; a.mask_kind = mask_kind;
    a.desc = reinterpret_cast<const SeqDesc*>(d_desc); a.key_mask = d_key_mask;
    a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind; a.lse = d_lse;
    a.drop = make_drop(drop_seed, drop_p);
    UNIMM_CHECK(drop_p <= 0.f || static_cast<double>(B) * heads * Sq * Skv < 4294967296.0, "attention dropout: 32-bit element index");
    return attention_mma_lp(a, static_cast<cudaStream_t>(stream));
}

size_t unimm_k_attention_backward_scratch(int B, int heads, int D, int Sq) { return attention_backward_scratch(B, heads, D, Sq); }

int unimm_k_attention_backward(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, const void* d_o, int ldo,
                               const float* d_dO, const float* d_lse, int B, int heads, int D, int Sq, int Skv, int mask_kind,
                               const unimm_seq_desc_t* d_desc, const float* d_key_mask, int lp_kind, float* d_dq, int lddq, float* d_dk,
                               int lddk, float* d_dv, int lddv, float* d_amax_accum, const float* d_dO_amax, uint32_t drop_seed, float drop_p,
                               void* d_scratch, size_t scratch_bytes, void* stream) {
    UNIMM_CHECK(lp_kind == LP_BF16 || lp_kind == LP_FP16, "attention backward: 16-bit tensors only");
    AttnArgs a;
    a.q = d_q; a.ldq = ldq; a.k = d_k; a.ldk = ldk; a.v = d_v; a.ldv = ldv; a.o = const_cast<void*>(d_o); a.ldo = ldo;
    a.B = B; a.heads = heads; a.D = D; a.Sq = Sq; a.Skv = Skv; a.mask_kind = mask_kind;
    a.desc = reinterpret_cast<const SeqDesc*>(d_desc); a.key_mask = d_key_mask;
    a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind;
    a.drop = make_drop(drop_seed, drop_p);
    return attention_backward_lp(a, d_dO, heads * D, d_lse, d_dq, lddq, d_dk, lddk, d_dv, lddv, d_scratch, scratch_bytes,
                                 static_cast<cudaStream_t>(stream), d_amax_accum, d_dO_amax);
}

int unimm_k_layernorm_backward_amax(const float* d_dy, const float* d_x, int rows, int H, const float* d_gamma, float* d_dx, float* d_dgamma,
                                    float* d_dbeta, float* d_amax, void* stream) {
    UNIMM_CHECK(d_dy && d_x && d_gamma && d_dx && d_dgamma && d_dbeta, "null argument");
    return layernorm_backward(d_dy, d_x, rows, H, d_gamma, d_dx, d_dgamma, d_dbeta, static_cast<cudaStream_t>(stream), d_amax);
}

int unimm_k_gelu_backward_amax(const float* d_dy, const float* d_x, int64_t n, float* d_dx, float* d_amax, void* stream) {
    UNIMM_CHECK(d_dy && d_x && d_dx && n > 0, "bad argument");
    return gelu_backward(d_dy, d_x, static_cast<size_t>(n), d_dx, static_cast<cudaStream_t>(stream), d_amax);
}

}  // extern "C"
