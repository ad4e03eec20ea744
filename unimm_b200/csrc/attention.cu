// Fused masked attention, CUDA-core (SIMT) flavour: softmax(Q K^T / sqrt(d) + mask) V with the mask
// generated on the fly from the 4-integer sequence descriptor instead of the reference's dense
// additive [B,1,S,S] fp32 tensor (reference models/vilbert_dialog.py:395-410 text, :524-539 image,
// :681-721 co-attention; mask construction :1396-1431 and utils/data_utils.py:149-210, :353-354).
//
// Flash-style: one CTA owns 32 query rows of one (sequence, head); keys/values stream through shared
// memory in tiles of 64 with an online softmax, and key tiles outside the union of the rows' allowed
// intervals are never loaded.  Used by the fp32 parity mode (T = float) and as the reference
// implementation for the tensor-core attention (T = bf16 inputs, fp32 math).
//
// Masked entries are excluded exactly (probability 0).  The reference adds -10000 instead; for any row
// with at least one allowed key exp(-10000 + ...) underflows to exactly 0 in fp32, so the results are
// identical.  Rows with NO allowed key (padding rows) get -10000 on every column in the reference,
// i.e. a softmax over the raw scores: those rows attend to all keys here as well.
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

constexpr int QT = 32;   // query rows per CTA
constexpr int KT = 64;   // keys per shared-memory tile
constexpr int NW = 8;    // warps per CTA
constexpr int RPW = QT / NW;  // query rows per warp

// fp32 parity mode uses the accurate expf; the bf16 instantiation the SFU ex2 path
template <typename T>
__device__ __forceinline__ float attn_exp(float x);
template <>
__device__ __forceinline__ float attn_exp<float>(float x) { return expf(x); }
template <>
__device__ __forceinline__ float attn_exp<bf16>(float x) { return __expf(x); }
template <>
__device__ __forceinline__ float attn_exp<fp16>(float x) { return __expf(x); }

template <typename T, int D>
__global__ void __launch_bounds__(NW * 32)
attn_simt_kernel(AttnArgs a) {
    extern __shared__ float smem[];
    float* Qs = smem;                       // [QT][D]
    float* Ks = Qs + QT * D;                // [KT][D+1]
    float* Vs = Ks + KT * (D + 1);          // [KT][D]
    float* Ps = Vs + KT * D;                // [NW][RPW][KT]
    __shared__ int s_any_key;

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Sq = a.Sq, Skv = a.Skv;
    const T* Q = static_cast<const T*>(a.q) + static_cast<size_t>(b) * Sq * a.ldq + h * D;
    const T* K = static_cast<const T*>(a.k) + static_cast<size_t>(b) * Skv * a.ldk + h * D;
    const T* V = static_cast<const T*>(a.v) + static_cast<size_t>(b) * Skv * a.ldv + h * D;
    T* O = static_cast<T*>(a.o) + static_cast<size_t>(b) * Sq * a.ldo + h * D;

    // ---- per-row allowed sets -------------------------------------------------------------------
    SeqDesc desc = {0, 0, 0, 0};
    if (a.mask_kind != MASK_KEY_VECTOR) desc = a.desc[b];
    const float* kmask = a.mask_kind == MASK_KEY_VECTOR ? a.key_mask + static_cast<size_t>(b) * Skv : nullptr;
    if (tid == 0) s_any_key = 0;
    __syncthreads();
    if (kmask != nullptr) {
        int any = 0;
        for (int k = tid; k < Skv; k += blockDim.x) any |= (kmask[k] > 0.5f);
        if (any) atomicOr(&s_any_key, 1);
    }
    // Q tile -> smem (fp32)
    for (int i = tid; i < QT * D; i += blockDim.x) {
        const int r = i / D, c = i % D;
        Qs[i] = (q0 + r < Sq) ? to_f32<T>(Q[static_cast<size_t>(q0 + r) * a.ldq + c]) : 0.f;
    }
    __syncthreads();
    const bool key_mask_all = (kmask != nullptr) && (s_any_key == 0);  // every key masked: uniform shift only

    int lo[RPW], hi[RPW], self[RPW];
    int kv_lo = Skv, kv_hi = 0;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        lo[r] = 0; hi[r] = Skv; self[r] = -1;
    }
    if (a.mask_kind == MASK_TEXT_SELF) {
        // union over ALL rows of the CTA (tile loads are CTA-wide), individual sets per owned row
        for (int r = 0; r < QT; ++r) {
            const int qr = q0 + r;
            if (qr >= Sq) break;
            int l, hh, s;
            text_row_interval(desc, qr, Skv, l, hh, s);
            if (hh <= l && s < 0) { l = 0; hh = Skv; }  // padding row: all keys
            kv_lo = min(kv_lo, l);
            kv_hi = max(kv_hi, max(hh, s + 1));
            if (r / RPW == warp) { lo[r % RPW] = l; hi[r % RPW] = hh; self[r % RPW] = s; }
        }
    } else if (a.mask_kind == MASK_CO_INTERVAL) {
        int l, hh;
        co_interval(desc, Skv, l, hh);
        if (hh <= l) { l = 0; hh = Skv; }
#pragma unroll
        for (int r = 0; r < RPW; ++r) { lo[r] = l; hi[r] = hh; }
        kv_lo = l; kv_hi = hh;
    } else {
        kv_lo = 0; kv_hi = Skv;
    }
    // rows are laid out warp-major: warp w owns tile rows w*RPW .. w*RPW+RPW-1
    float m[RPW], l_sum[RPW], o[RPW][D / 32];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        m[r] = -INFINITY; l_sum[r] = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) o[r][i] = 0.f;
    }
    float* Pw = Ps + warp * RPW * KT;

    const int t_begin = (kv_lo / KT) * KT;
    for (int t0 = t_begin; t0 < kv_hi; t0 += KT) {
        __syncthreads();  // previous tile fully consumed
        for (int i = tid; i < KT * D; i += blockDim.x) {
            const int kr = i / D, c = i % D;
            const int key = t0 + kr;
            float kv = 0.f, vv = 0.f;
            if (key < Skv) {
                kv = to_f32<T>(K[static_cast<size_t>(key) * a.ldk + c]);
                vv = to_f32<T>(V[static_cast<size_t>(key) * a.ldv + c]);
            }
            Ks[kr * (D + 1) + c] = kv;
            Vs[kr * D + c] = vv;
        }
        __syncthreads();

        float s[RPW][2];
#pragma unroll
        for (int r = 0; r < RPW; ++r) s[r][0] = s[r][1] = 0.f;
        const float* k0p = Ks + lane * (D + 1);
        const float* k1p = Ks + (lane + 32) * (D + 1);
        const float* qp = Qs + warp * RPW * D;
#pragma unroll 8
        for (int d = 0; d < D; ++d) {
            const float k0 = k0p[d], k1 = k1p[d];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const float qv = qp[r * D + d];
                s[r][0] = fmaf(qv, k0, s[r][0]);
                s[r][1] = fmaf(qv, k1, s[r][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float p[2];
            float tmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int key = t0 + lane + 32 * j;
                bool ok = key < Skv;
                if (a.mask_kind == MASK_KEY_VECTOR) ok = ok && (key_mask_all || kmask[min(key, Skv - 1)] > 0.5f);
                else ok = ok && ((key >= lo[r] && key < hi[r]) || key == self[r]);
                s[r][j] = ok ? s[r][j] * a.scale : -INFINITY;
                tmax = fmaxf(tmax, s[r][j]);
            }
            tmax = warp_max(tmax);
            const float m_new = fmaxf(m[r], tmax);
            float corr = 1.f;
            if (m_new == -INFINITY) {
                p[0] = p[1] = 0.f;
            } else {
                corr = attn_exp<T>(m[r] - m_new);  // m[r] = -inf -> 0
                p[0] = attn_exp<T>(s[r][0] - m_new);
                p[1] = attn_exp<T>(s[r][1] - m_new);
            }
            l_sum[r] = l_sum[r] * corr + warp_sum(p[0] + p[1]);
            m[r] = m_new;
#pragma unroll
            for (int i = 0; i < D / 32; ++i) o[r][i] *= corr;
            Pw[r * KT + lane] = p[0];
            Pw[r * KT + lane + 32] = p[1];
        }
        __syncwarp();
#pragma unroll 4
        for (int key = 0; key < KT; ++key) {
            float vv[D / 32];
#pragma unroll
            for (int i = 0; i < D / 32; ++i) vv[i] = Vs[key * D + lane + 32 * i];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const float pv = Pw[r * KT + key];
#pragma unroll
                for (int i = 0; i < D / 32; ++i) o[r][i] = fmaf(pv, vv[i], o[r][i]);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int qr = q0 + warp * RPW + r;
        if (qr >= Sq) continue;
        const float inv = l_sum[r] > 0.f ? 1.0f / l_sum[r] : 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i)
            O[static_cast<size_t>(qr) * a.ldo + lane + 32 * i] = from_f32<T>(o[r][i] * inv);
    }
}

template <typename T, int D>
int launch_simt(const AttnArgs& a, cudaStream_t stream) {
    const size_t smem = sizeof(float) * (QT * D + KT * (D + 1) + KT * D + NW * RPW * KT);
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_simt_kernel<T, D>), smem));
    dim3 grid((a.Sq + QT - 1) / QT, a.heads, a.B);
    UNIMM_CHECK(a.B <= 65535, "attention: batch too large for one launch");
    attn_simt_kernel<T, D><<<grid, NW * 32, smem, stream>>>(a);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int check_args(const AttnArgs& a) {
    UNIMM_CHECK(a.B > 0 && a.heads > 0 && a.Sq > 0 && a.Skv > 0, "attention: empty problem");
    UNIMM_CHECK(a.D == 64 || a.D == 128, "attention: head dim must be 64 or 128");
    UNIMM_CHECK(a.mask_kind == MASK_KEY_VECTOR ? a.key_mask != nullptr : a.desc != nullptr, "attention: mask operand missing");
    return 0;
}

}  // namespace

int attention_simt_f32(const AttnArgs& a, cudaStream_t stream) {
    UNIMM_TRY(check_args(a));
    return a.D == 64 ? launch_simt<float, 64>(a, stream) : launch_simt<float, 128>(a, stream);
}

int attention_simt_lp(const AttnArgs& a, cudaStream_t stream) {
    UNIMM_TRY(check_args(a));
    if (a.lp_kind == LP_FP16) return a.D == 64 ? launch_simt<fp16, 64>(a, stream) : launch_simt<fp16, 128>(a, stream);
    return a.D == 64 ? launch_simt<bf16, 64>(a, stream) : launch_simt<bf16, 128>(a, stream);
}

}  // namespace unimm
