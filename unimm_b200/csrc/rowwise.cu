// HBM-bound row kernels: embedding gather + LayerNorm, LayerNorm, feature/row gathers, casts.
// One warp per row, 128-bit lane-interleaved accesses, warp-shuffle reductions, fp32 statistics.
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

constexpr float kLnEps = 1e-12f;  // BertLayerNorm eps (reference models/vilbert_dialog.py:322)

template <int NV>
__device__ __forceinline__ void ln_normalise_store(float4 (&x)[NV], int lane, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float* y_f32, bf16* y_bf16, int lp_kind) {
    constexpr int H = NV * 128;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + kLnEps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        float4 y;
        y.x = (x[i].x - mean) * rstd * g.x + b.x;
        y.y = (x[i].y - mean) * rstd * g.y + b.y;
        y.z = (x[i].z - mean) * rstd * g.z + b.z;
        y.w = (x[i].w - mean) * rstd * g.w + b.w;
        if (y_f32 != nullptr) *reinterpret_cast<float4*>(y_f32 + c) = y;
        if (y_bf16 != nullptr && lp_kind == LP_HILO) {      // fp32-class mode: hi plane at [0, H), lo plane at [H, 2H) of the row
            uint2 hi, lo;
            split_hilo2(y.x, y.y, hi.x, lo.x);
            split_hilo2(y.z, y.w, hi.y, lo.y);
            *reinterpret_cast<uint2*>(y_bf16 + c) = hi;
            *reinterpret_cast<uint2*>(y_bf16 + H + c) = lo;
        } else if (y_bf16 != nullptr) {
            uint2 p;
            p.x = pack_lp2(y.x, y.y, lp_kind);
            p.y = pack_lp2(y.z, y.w, lp_kind);
            *reinterpret_cast<uint2*>(y_bf16 + c) = p;
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(128)
layernorm_kernel(const float* __restrict__ x, int ldx, int rows, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* y_f32, bf16* y_bf16, int lp_kind) {
    constexpr int H = NV * 128;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<size_t>(row) * ldx;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xr + (lane + 32 * i) * 4);
    ln_normalise_store<NV>(v, lane, gamma, beta, y_f32 ? y_f32 + static_cast<size_t>(row) * H : nullptr,
                           y_bf16 ? y_bf16 + static_cast<size_t>(row) * (lp_kind == LP_HILO ? 2 * H : H) : nullptr, lp_kind);
}

template <int NV, typename IdT>
__global__ void __launch_bounds__(128)
embed_text_ln_kernel(const IdT* __restrict__ ids, const IdT* __restrict__ type_ids,
                     const IdT* __restrict__ pos_ids, int rows, int vocab, int max_pos, int type_vocab, int type_ext,
                     const float* __restrict__ word_emb, const float* __restrict__ pos_emb,
                     const float* __restrict__ type_emb, const float* __restrict__ type_ext_emb,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float* out_f32, bf16* out_bf16,
                     int lp_kind, int* err_flag) {
    constexpr int H = NV * 128;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    long long id = ids[row], ty = type_ids[row], pos = pos_ids[row];
    if (id < 0 || id >= vocab || pos < 0 || pos >= max_pos || ty < 0 || ty >= type_vocab + type_ext) {
        if (lane == 0) atomicExch(err_flag, 1);  // the reference would raise an index error here
        id = min(max(id, 0LL), (long long)vocab - 1);
        pos = min(max(pos, 0LL), (long long)max_pos - 1);
        ty = min(max(ty, 0LL), (long long)(type_vocab + type_ext - 1));
    }
    const float* w = word_emb + static_cast<size_t>(id) * H;
    const float* p = pos_emb + static_cast<size_t>(pos) * H;
    // ids >= type_vocab_size select the extension table (reference :337-350)
    const float* t = (ty < type_vocab) ? type_emb + static_cast<size_t>(ty) * H
                                       : type_ext_emb + static_cast<size_t>(ty - type_vocab) * H;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        const float4 a = __ldg(reinterpret_cast<const float4*>(w + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + c));
        const float4 d = __ldg(reinterpret_cast<const float4*>(t + c));
        v[i] = make_float4((a.x + b.x) + d.x, (a.y + b.y) + d.y, (a.z + b.z) + d.z, (a.w + b.w) + d.w);
    }
    ln_normalise_store<NV>(v, lane, gamma, beta, out_f32 ? out_f32 + static_cast<size_t>(row) * H : nullptr,
                           out_bf16 ? out_bf16 + static_cast<size_t>(row) * (lp_kind == LP_HILO ? 2 * H : H) : nullptr, lp_kind);
}

__global__ void image_loc_kernel(const float* __restrict__ loc, const int* __restrict__ feat_index, int R, int H,
                                 const float* __restrict__ Wloc, const float* __restrict__ bloc, float* __restrict__ out) {
    const int row = blockIdx.x;  // b*R + r
    const int b = row / R, r = row % R;
    const int src = (feat_index ? feat_index[b] : b) * R + r;
    float l[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) l[j] = __ldg(loc + static_cast<size_t>(src) * 5 + j);
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) acc = fmaf(l[j], __ldg(Wloc + h * 5 + j), acc);
        out[static_cast<size_t>(row) * H + h] = acc + __ldg(bloc + h);
    }
}

__global__ void gather_features_kernel(const float* __restrict__ feat, const int* __restrict__ feat_index, int R, int F,
                                       float* dst_f32, bf16* dst_bf16, int lp_kind) {
    const int row = blockIdx.x;
    const int b = row / R, r = row % R;
    const float* s = feat + (static_cast<size_t>(feat_index ? feat_index[b] : b) * R + r) * F;
    for (int c = threadIdx.x * 4; c < F; c += blockDim.x * 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s + c));
        if (dst_f32) *reinterpret_cast<float4*>(dst_f32 + static_cast<size_t>(row) * F + c) = v;
        if (dst_bf16) {
            uint2 p;
            p.x = pack_lp2(v.x, v.y, lp_kind);
            p.y = pack_lp2(v.z, v.w, lp_kind);
            *reinterpret_cast<uint2*>(dst_bf16 + static_cast<size_t>(row) * F + c) = p;
        }
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ src_f32, const bf16* __restrict__ src_bf16,
                                   const int* __restrict__ rows, int H, float* dst_f32, bf16* dst_bf16) {
    const int i = blockIdx.x;
    const size_t s = static_cast<size_t>(rows[i]) * H, d = static_cast<size_t>(i) * H;
    for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
        if (dst_f32) *reinterpret_cast<float4*>(dst_f32 + d + c) = *reinterpret_cast<const float4*>(src_f32 + s + c);
        if (dst_bf16) *reinterpret_cast<uint2*>(dst_bf16 + d + c) = *reinterpret_cast<const uint2*>(src_bf16 + s + c);
    }
}

__global__ void cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n4, int lp_kind) {
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        uint2 p;
        p.x = pack_lp2(v.x, v.y, lp_kind);
        p.y = pack_lp2(v.z, v.w, lp_kind);
        reinterpret_cast<uint2*>(dst)[i] = p;
    }
}

__global__ void gather_labels_kernel(const int64_t* __restrict__ labels, const int* __restrict__ rows, int n,
                                     int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<int>(labels[rows[i]]);
}

__global__ void expand_key_mask_kernel(const float* __restrict__ mask, const int* __restrict__ index, int B, int R,
                                       float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * R) out[i] = mask[static_cast<size_t>(index[i / R]) * R + (i % R)];
}

}  // namespace

int gather_labels(const int64_t* labels, const int* rows, int n, int* out, cudaStream_t stream) {
    if (n == 0) return 0;
    gather_labels_kernel<<<(n + 255) / 256, 256, 0, stream>>>(labels, rows, n, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int expand_key_mask(const float* mask, const int* index, int B, int R, float* out, cudaStream_t stream) {
    expand_key_mask_kernel<<<(B * R + 255) / 256, 256, 0, stream>>>(mask, index, B, R, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int layernorm_rows(const float* x, int ldx, int rows, int H, const float* gamma, const float* beta, float* y_f32,
                   bf16* y_bf16, int lp_kind, cudaStream_t stream) {
    UNIMM_CHECK(rows > 0, "layernorm: no rows");
    UNIMM_CHECK((ldx & 3) == 0, "layernorm: ldx must be a multiple of 4");
    const int grid = (rows + 3) / 4;
    if (H == 768) layernorm_kernel<6><<<grid, 128, 0, stream>>>(x, ldx, rows, gamma, beta, y_f32, y_bf16, lp_kind);
    else if (H == 1024) layernorm_kernel<8><<<grid, 128, 0, stream>>>(x, ldx, rows, gamma, beta, y_f32, y_bf16, lp_kind);
    else UNIMM_CHECK(false, "layernorm: hidden size must be 768 or 1024");
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

template <typename IdT>
static int embed_text_ln_t(const IdT* ids, const IdT* type_ids, const IdT* pos_ids, int rows, int H, int vocab, int max_pos,
                           int type_vocab, int type_ext, const float* word_emb, const float* pos_emb, const float* type_emb,
                           const float* type_ext_emb, const float* gamma, const float* beta, float* out_f32, bf16* out_bf16,
                           int lp_kind, int* err_flag, cudaStream_t stream) {
    UNIMM_CHECK(rows > 0, "embed: no rows");
    const int grid = (rows + 3) / 4;
    if (H == 768)
        embed_text_ln_kernel<6, IdT><<<grid, 128, 0, stream>>>(ids, type_ids, pos_ids, rows, vocab, max_pos, type_vocab, type_ext, word_emb,
                                                               pos_emb, type_emb, type_ext_emb, gamma, beta, out_f32, out_bf16, lp_kind, err_flag);
    else if (H == 1024)
        embed_text_ln_kernel<8, IdT><<<grid, 128, 0, stream>>>(ids, type_ids, pos_ids, rows, vocab, max_pos, type_vocab, type_ext, word_emb,
                                                               pos_emb, type_emb, type_ext_emb, gamma, beta, out_f32, out_bf16, lp_kind, err_flag);
    else UNIMM_CHECK(false, "embed: hidden size must be 768 or 1024");
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int embed_text_ln(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int rows, int H, int vocab,
                  int max_pos, int type_vocab, int type_ext, const float* word_emb, const float* pos_emb,
                  const float* type_emb, const float* type_ext_emb, const float* gamma, const float* beta, float* out_f32,
                  bf16* out_bf16, int lp_kind, int* err_flag, cudaStream_t stream) {
    return embed_text_ln_t<int64_t>(ids, type_ids, pos_ids, rows, H, vocab, max_pos, type_vocab, type_ext, word_emb, pos_emb, type_emb,
                                    type_ext_emb, gamma, beta, out_f32, out_bf16, lp_kind, err_flag, stream);
}

int embed_text_ln_i32(const int32_t* ids, const int32_t* type_ids, const int32_t* pos_ids, int rows, int H, int vocab,
                      int max_pos, int type_vocab, int type_ext, const float* word_emb, const float* pos_emb,
                      const float* type_emb, const float* type_ext_emb, const float* gamma, const float* beta, float* out_f32,
                      bf16* out_bf16, int lp_kind, int* err_flag, cudaStream_t stream) {
    return embed_text_ln_t<int32_t>(ids, type_ids, pos_ids, rows, H, vocab, max_pos, type_vocab, type_ext, word_emb, pos_emb, type_emb,
                                    type_ext_emb, gamma, beta, out_f32, out_bf16, lp_kind, err_flag, stream);
}

int image_loc_embed(const float* loc, const int* feat_index, int B, int R, int H, const float* Wloc, const float* bloc,
                    float* out, cudaStream_t stream) {
    image_loc_kernel<<<B * R, 256, 0, stream>>>(loc, feat_index, R, H, Wloc, bloc, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int gather_features(const float* feat, const int* feat_index, int B, int R, int F, float* dst_f32, bf16* dst_bf16,
                    int lp_kind, cudaStream_t stream) {
    UNIMM_CHECK((F & 3) == 0, "feature size must be a multiple of 4");
    gather_features_kernel<<<B * R, 256, 0, stream>>>(feat, feat_index, R, F, dst_f32, dst_bf16, lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int gather_rows(const float* src_f32, const bf16* src_bf16, const int* rows, int n, int H, float* dst_f32, bf16* dst_bf16,
                cudaStream_t stream) {
    if (n == 0) return 0;
    UNIMM_CHECK((H & 3) == 0, "gather_rows: H must be a multiple of 4");
    gather_rows_kernel<<<n, 192, 0, stream>>>(src_f32, src_bf16, rows, H, dst_f32, dst_bf16);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

// 16-bit activation copy -> fp32 (refreshes the fp32 view of a residual stream that was kept in 16 bits)
__global__ void widen_kernel(const uint2* __restrict__ src, float4* __restrict__ dst, size_t n4, int lp_kind) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint2 u = src[i];
        float4 o;
        if (lp_kind == LP_FP16) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            o = make_float4(a.x, a.y, b.x, b.y);
        } else {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)),
                         b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            o = make_float4(a.x, a.y, b.x, b.y);
        }
        dst[i] = o;
    }
}

int cast_lp_to_f32(const bf16* src, float* dst, size_t n, int lp_kind, cudaStream_t stream) {
    UNIMM_CHECK((n & 3) == 0, "cast: element count must be a multiple of 4");
    const size_t n4 = n / 4;
    int grid = static_cast<int>((n4 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    widen_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint2*>(src), reinterpret_cast<float4*>(dst), n4, lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

// fp32 [rows, K] (leading dimension ldx) -> the two fp16 planes [rows, 2K] of the fp32-class mode (LP_HILO).
// Source row of output row r: row_idx ? row_idx[r] : r.
__global__ void split_hilo_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ row_idx, int rows, int K, bf16* __restrict__ out) {
    const int kv = K / 4;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < static_cast<size_t>(rows) * kv;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / kv;
        const int c = static_cast<int>(i % kv) * 4;
        const size_t sr = row_idx ? static_cast<size_t>(row_idx[r]) : r;
        const float4 v = *reinterpret_cast<const float4*>(x + sr * ldx + c);
        uint2 hi, lo;
        split_hilo2(v.x, v.y, hi.x, lo.x);
        split_hilo2(v.z, v.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(out + r * 2 * K + c) = hi;
        *reinterpret_cast<uint2*>(out + r * 2 * K + K + c) = lo;
    }
}
int split_f32_to_hilo(const float* x, int ldx, int rows, int K, bf16* out, cudaStream_t stream, const int* row_idx) {
    UNIMM_CHECK(rows > 0 && K % 4 == 0 && ldx % 4 == 0, "split: K and ldx must be multiples of 4");
    const size_t n = static_cast<size_t>(rows) * (K / 4);
    int grid = static_cast<int>((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    split_hilo_kernel<<<grid, 256, 0, stream>>>(x, ldx, row_idx, rows, K, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

// ---- backward of y = LayerNorm(x) * gamma + beta (eps inside the sqrt, biased variance — BertLayerNorm, reference :270-279) ----
//   xhat = (x - mean) * rstd;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma = sum_rows dy * xhat;  dbeta = sum_rows dy
// One warp per row (grid-stride); every lane keeps the column sums of ITS columns in registers over all the rows its warp handles and
// adds them to the per-block shared accumulators once at the end (one atomic per column, warp and block instead of one per element).
template <int NV>
__global__ void __launch_bounds__(128)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int rows, const float* __restrict__ gamma,
                     float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, unsigned* __restrict__ amax) {
    constexpr int H = NV * 128;
    __shared__ float s_dg[H], s_db[H];
    for (int i = threadIdx.x; i < H; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = 0.f;
    float4 acc_g[NV], acc_b[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc_g[i] = acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int row = blockIdx.x * 4 + warp; row < rows; row += gridDim.x * 4) {
        float4 xv[NV], gv[NV], dv[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            xv[i] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * H + (lane + 32 * i) * 4);
            dv[i] = *reinterpret_cast<const float4*>(dy + static_cast<size_t>(row) * H + (lane + 32 * i) * 4);
            s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
        }
        const float mean = warp_sum(s) * (1.0f / H);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
            q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + kLnEps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
            xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;            // xhat
            gv[i] = make_float4(dv[i].x * gm.x, dv[i].y * gm.y, dv[i].z * gm.z, dv[i].w * gm.w);
            sg += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
            sgx += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
        }
        const float mg = warp_sum(sg) * (1.0f / H), mgx = warp_sum(sgx) * (1.0f / H);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            float4 o;
            o.x = rstd * (gv[i].x - mg - xv[i].x * mgx); o.y = rstd * (gv[i].y - mg - xv[i].y * mgx);
            o.z = rstd * (gv[i].z - mg - xv[i].z * mgx); o.w = rstd * (gv[i].w - mg - xv[i].w * mgx);
            *reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * H + c) = o;
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
            acc_g[i].x += dv[i].x * xv[i].x; acc_g[i].y += dv[i].y * xv[i].y; acc_g[i].z += dv[i].z * xv[i].z; acc_g[i].w += dv[i].w * xv[i].w;
            acc_b[i].x += dv[i].x; acc_b[i].y += dv[i].y; acc_b[i].z += dv[i].z; acc_b[i].w += dv[i].w;
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        atomicAdd(&s_dg[c], acc_g[i].x); atomicAdd(&s_dg[c + 1], acc_g[i].y); atomicAdd(&s_dg[c + 2], acc_g[i].z); atomicAdd(&s_dg[c + 3], acc_g[i].w);
        atomicAdd(&s_db[c], acc_b[i].x); atomicAdd(&s_db[c + 1], acc_b[i].y); atomicAdd(&s_db[c + 2], acc_b[i].z); atomicAdd(&s_db[c + 3], acc_b[i].w);
    }
    if (amax != nullptr) {
        const unsigned u = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
        if (lane == 0) atomicMax(amax, u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) { atomicAdd(dgamma + i, s_dg[i]); atomicAdd(dbeta + i, s_db[i]); }
}
int layernorm_backward(const float* dy, const float* x, int rows, int H, const float* gamma, float* dx, float* dgamma, float* dbeta,
                       cudaStream_t stream, float* amax_out) {
    UNIMM_CHECK(rows > 0, "layernorm backward: no rows");
    unsigned* am = reinterpret_cast<unsigned*>(amax_out);
    if (am != nullptr) UNIMM_CUDA_CHECK(cudaMemsetAsync(am, 0, sizeof(unsigned), stream));
    UNIMM_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * H, stream));
    UNIMM_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * H, stream));
    int grid = (rows + 3) / 4;
    if (grid > 148 * 4) grid = 148 * 4;
    if (H == 768) layernorm_bwd_kernel<6><<<grid, 128, 0, stream>>>(dy, x, rows, gamma, dx, dgamma, dbeta, am);
    else if (H == 1024) layernorm_bwd_kernel<8><<<grid, 128, 0, stream>>>(dy, x, rows, gamma, dx, dgamma, dbeta, am);
    else UNIMM_CHECK(false, "layernorm backward: hidden size must be 768 or 1024");
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

// backward of the exact (erf) GELU (reference :115-121): dx = dy * (Phi(x) + x * phi(x)); in place allowed (dx == dy)
__device__ __forceinline__ float gelu_grad(float v) {
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    return cdf + v * pdf;
}
// dy may alias dx (no __restrict__ on them); 128-bit accesses on the bulk, scalars on the tail
__global__ void gelu_bwd_kernel(const float* dy, const float* __restrict__ x, size_t n, float* dx, unsigned* __restrict__ amax) {
    float mx = 0.f;
    const size_t n4 = n / 4;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 d = reinterpret_cast<const float4*>(dy)[i];
        d.x *= gelu_grad(v.x); d.y *= gelu_grad(v.y); d.z *= gelu_grad(v.z); d.w *= gelu_grad(v.w);
        reinterpret_cast<float4*>(dx)[i] = d;
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(d.x), fabsf(d.y))), fmaxf(fabsf(d.z), fabsf(d.w)));
    }
    for (size_t i = n4 * 4 + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float d = dy[i] * gelu_grad(x[i]);
        dx[i] = d;
        mx = fmaxf(mx, fabsf(d));
    }
    if (amax != nullptr) {
        const unsigned u = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
        if ((threadIdx.x & 31) == 0) atomicMax(amax, u);
    }
}
int gelu_backward(const float* dy, const float* x, size_t n, float* dx, cudaStream_t stream, float* amax_out) {
    int grid = static_cast<int>((n / 4 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    unsigned* am = reinterpret_cast<unsigned*>(amax_out);
    if (am != nullptr) UNIMM_CUDA_CHECK(cudaMemsetAsync(am, 0, sizeof(unsigned), stream));
    UNIMM_CHECK((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0,
                "gelu backward: 16-byte aligned buffers");
    gelu_bwd_kernel<<<grid < 1 ? 1 : grid, 256, 0, stream>>>(dy, x, n, dx, am);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

__global__ void differ_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, int* flag) {
    bool d = false;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        d |= __float_as_uint(a[i]) != __float_as_uint(b[i]);
    if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}
int buffers_differ(const float* a, const float* b, size_t n, int* d_flag, cudaStream_t stream) {
    UNIMM_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, sizeof(int), stream));
    differ_kernel<<<148 * 8, 256, 0, stream>>>(a, b, n, d_flag);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int cast_f32_to_lp(const float* src, bf16* dst, size_t n, int lp_kind, cudaStream_t stream) {
    UNIMM_CHECK((n & 3) == 0, "cast: element count must be a multiple of 4");
    const size_t n4 = n / 4;
    int grid = static_cast<int>((n4 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    cast_kernel<<<grid, 256, 0, stream>>>(src, dst, n4, lp_kind);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace unimm
