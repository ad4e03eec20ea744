// Backward of the fused masked attention (dense [B, S] layout) on the tensor cores — SURVEY.md §8f item 1.
//
// Reference: the autograd of  softmax(Q K^T / sqrt(d) + mask) V  (models/vilbert_dialog.py:395-410 text / image self-attention,
// :681-721 the two co-attentions) under train.py:453's backward.  Forward = attention_mma.cu with AttnArgs::lse set; this file
// recomputes the probabilities from the saved row log-sum-exp instead of keeping a [B, heads, S, S] tensor:
//
//     P = exp(scale * Q K^T - lse)   (0 outside the row's allowed key set — the same sets as the forward, rebuilt from the
//                                     4-int descriptor / the key vector; there is no mask tensor)
//     dP = dO V^T,   delta = rowsum(dO o O),   dS = P o (dP - delta)
//     dQ = scale * dS K,    dK = scale * dS^T Q,    dV = P^T dO
//
// Two kernels, both FlashAttention-2 style with mma.sync.m16n8k16 and fp32 accumulators, neither uses atomics:
//   attn_bwd_dq_kernel   a warp owns 16 QUERY rows, loops over 64-key tiles:   S, dP -> dS -> dQ += dS K
//   attn_bwd_dkv_kernel  a warp owns 16 KEY rows, loops over 64-query tiles and works on the TRANSPOSED tiles
//                        S^T = K Q^T, dP^T = V dO^T, so that P^T / dS^T come out of the accumulators already in the A-fragment
//                        layout of dV += P^T dO and dK += dS^T Q
// dO arrives as fp32 (the out-projection's dgrad), is turned into a 16-bit operand with a power-of-two scale taken from its own
// maximum on the device (fp16: gradients sit far below the normal range), and the scale is divided out again when dQ / dK / dV are
// written as fp32.  dS is additionally shifted by 2^-4 in fp16 so that its largest entries stay below 65504.
#include "attn_common.cuh"
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

using namespace attn;

constexpr float kLog2e = 1.4426950408889634f;

struct BwdArgs {
    const bf16* q; int ldq;
    const bf16* k; int ldk;
    const bf16* v; int ldv;
    const bf16* dO; int lddo;       // scaled 16-bit copy of the incoming gradient
    const float* lse;               // [B, heads, Sq]
    const float* delta;             // [B, heads, Sq]  rowsum(dO16 o O16)
    float* dq; int lddq;
    float* dk; int lddk;
    float* dv; int lddv;
    int B, heads, Sq, Skv, mask_kind;
    const SeqDesc* desc;
    const float* key_mask;
    float scale;
    const float* inv_scale;         // device: 1 / (scale applied to dO)
    float ds_shift;                 // factor applied to dS before it is rounded to 16 bits
    unsigned* amax;                 // optional: max |dq|, |dk|, |dv| written (float bits, atomicMax) for the consumer's 16-bit scale
    DropArgs drop;                  // dropout on the probabilities in the forward: O = (P o mask / (1 - p)) V, so dP = (dO V^T) o mask / (1 - p),
                                    //   dV = (P o mask / (1 - p))^T dO, and delta = rowsum(dO o O) as before
};

// allowed key set of query row qr: [lo, hi) U {self}, the forward's row_set (attention_mma.cu) — padding rows included, so that
// the recomputed P matches the saved lse
__device__ __forceinline__ void row_set(int mask_kind, const SeqDesc& desc, int Skv, int qr, int& lo, int& hi, int& self) {
    self = -1;
    if (mask_kind == MASK_TEXT_SELF) {
        text_row_interval(desc, qr, Skv, lo, hi, self);
        if (hi <= lo && self < 0) { lo = 0; hi = 1; }
        hi = min(hi, Skv);
    } else if (mask_kind == MASK_CO_INTERVAL) {
        co_interval(desc, Skv, lo, hi);
        if (hi <= lo) { lo = 0; hi = Skv; }
    } else {
        lo = 0; hi = Skv;
    }
}

// Rows beyond these bounds are padding: no real query attends such a key, and such a query's dO is exactly zero (nothing labelled can
// see a padding row, so no gradient ever reaches one — the backward of every row-wise operation maps a zero row to a zero row).  Whole
// tiles of them are skipped: their dQ / dK / dV are zeros.
__device__ __forceinline__ void real_extents(int mask_kind, const SeqDesc& d, int Sq, int Skv, int& q_real, int& k_real) {
    q_real = Sq; k_real = Skv;
    if (mask_kind == MASK_TEXT_SELF) {
        const int T = d.mode == 1 ? d.L : d.L + d.last_len;
        q_real = k_real = max(1, min(T, Skv));          // key 0 stays: the fallback of padding rows reads it (with a zero dO)
    } else if (mask_kind == MASK_CO_INTERVAL) {
        int lo, hi;
        co_interval(d, Skv, lo, hi);
        if (hi > lo) k_real = hi;
    }
}

// KEY_VECTOR: 64-bit words of allowed keys (all ones for the other kinds; "no valid key" = all keys, as the forward)
__device__ __forceinline__ void build_key_bits(const BwdArgs& a, int b, int kv_rows, unsigned long long* s_bits, int tid, int nthreads) {
    if (tid < 4) s_bits[tid] = 0ull;
    __syncthreads();
    if (a.mask_kind == MASK_KEY_VECTOR) {
        const float* km = a.key_mask + static_cast<size_t>(b) * a.Skv;
        for (int key = tid; key < kv_rows; key += nthreads)
            if (key < a.Skv && km[key] > 0.5f) atomicOr(&s_bits[key >> 6], 1ull << (key & 63));
        __syncthreads();
        if (tid == 0 && (s_bits[0] | s_bits[1] | s_bits[2] | s_bits[3]) == 0ull) s_bits[0] = s_bits[1] = s_bits[2] = s_bits[3] = ~0ull;
    } else {
        if (tid < 4) s_bits[tid] = ~0ull;
    }
    __syncthreads();
}

template <int D>
__device__ __forceinline__ void stage_rows(bf16* dst, const bf16* src, int ld, int r0, int n_stage, int n_valid_end, int tid, int nthreads) {
    constexpr int LD = D + PADE, CH = D / 8;
    for (int i = tid; i < n_stage * CH; i += nthreads) {
        const int r = i / CH, c = (i % CH) * 8;
        bf16* d = dst + r * LD + c;
        if (r0 + r < n_valid_end) cp_async16(d, src + static_cast<size_t>(r0 + r) * ld + c);
        else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
    }
}

// ------------------------------------------------------------------------------------------------ delta = rowsum(dO o O)
__global__ void attn_delta_kernel(const bf16* dO, int lddo, const bf16* O, int ldo, int rows_total, int Sq, int heads, int D, int lp_kind,
                                  float* delta) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows_total * heads) return;
    const int row = w / heads, h = w % heads;
    const bf16* a = dO + static_cast<size_t>(row) * lddo + h * D;
    const bf16* o = O + static_cast<size_t>(row) * ldo + h * D;
    float s = 0.f;
    for (int c = lane * 2; c < D; c += 64) {
        const uint32_t ua = *reinterpret_cast<const uint32_t*>(a + c), uo = *reinterpret_cast<const uint32_t*>(o + c);
        float2 fa, fo;
        if (lp_kind == LP_FP16) {
            fa = __half22float2(*reinterpret_cast<const __half2*>(&ua));
            fo = __half22float2(*reinterpret_cast<const __half2*>(&uo));
        } else {
            fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ua));
            fo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uo));
        }
        s += fa.x * fo.x + fa.y * fo.y;
    }
    s = warp_sum(s);
    const int b = row / Sq, r = row % Sq;
    if (lane == 0) delta[(static_cast<size_t>(b) * heads + h) * Sq + r] = s;
}

// ------------------------------------------------------------------------------------------------ dQ
template <int D, bool FP16, int NW>
__global__ void __launch_bounds__(NW * 32)
attn_bwd_dq_kernel(BwdArgs a, int kv_rows) {
    constexpr int MQT = 16 * NW, LD = D + PADE, NT = NW * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);             // [MQT][LD]
    bf16* Gs = Qs + MQT * LD;                                 // dO tile [MQT][LD]
    bf16* Ks = Gs + MQT * LD;                                 // [kv_rows][LD]
    bf16* Vs = Ks + static_cast<size_t>(kv_rows) * LD;        // [kv_rows][LD]
    __shared__ unsigned long long s_bits[4];

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MQT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int Sq = a.Sq, Skv = a.Skv;
    const bf16* Q = a.q + static_cast<size_t>(b) * Sq * a.ldq + h * D;
    const bf16* K = a.k + static_cast<size_t>(b) * Skv * a.ldk + h * D;
    const bf16* V = a.v + static_cast<size_t>(b) * Skv * a.ldv + h * D;
    const bf16* G = a.dO + static_cast<size_t>(b) * Sq * a.lddo + h * D;

    SeqDesc desc = {0, 0, 0, 0};
    if (a.mask_kind != MASK_KEY_VECTOR) desc = a.desc[b];
    int q_real, k_real;
    real_extents(a.mask_kind, desc, Sq, Skv, q_real, k_real);
    if (q0 >= q_real) {                       // a block of padding queries: zeros (block-uniform exit before any barrier)
        float* DQ = a.dq + static_cast<size_t>(b) * Sq * a.lddq + h * D;
        for (int i = tid; i < MQT * (D / 4); i += NT) {
            const int r = q0 + i / (D / 4), c = (i % (D / 4)) * 4;
            if (r < Sq) *reinterpret_cast<float4*>(DQ + static_cast<size_t>(r) * a.lddq + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    const int kv_stage = min(kv_rows, ((k_real + MKT - 1) / MKT) * MKT);      // rows beyond are never touched (w_end <= kv_stage)
    stage_rows<D>(Qs, Q, a.ldq, q0, MQT, Sq, tid, NT);
    stage_rows<D>(Gs, G, a.lddo, q0, MQT, Sq, tid, NT);
    stage_rows<D>(Ks, K, a.ldk, 0, kv_stage, Skv, tid, NT);
    stage_rows<D>(Vs, V, a.ldv, 0, kv_stage, Skv, tid, NT);
    build_key_bits(a, b, kv_rows, s_bits, tid, NT);
    cp_async_wait_all();
    __syncthreads();

    const int row[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
    int lo[2], hi[2], self[2];
    float lse2[2], dl[2];
    const size_t stat0 = (static_cast<size_t>(b) * a.heads + h) * Sq;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        row_set(a.mask_kind, desc, Skv, min(row[r], Sq - 1), lo[r], hi[r], self[r]);
        const bool ok = row[r] < Sq;
        lse2[r] = ok ? a.lse[stat0 + row[r]] * kLog2e : INFINITY;     // padding rows: exp2(-inf) = 0 everywhere
        dl[r] = ok ? a.delta[stat0 + row[r]] : 0.f;
    }
    // the warp's key range (rounded to tiles): tiles beyond it hold no allowed key for any of its rows
    int w_hi = max(max(hi[0], self[0] + 1), max(hi[1], self[1] + 1));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
    const int w_end = min(((w_hi + MKT - 1) / MKT) * MKT, kv_stage);

    const float sl = a.scale * kLog2e;
    const uint32_t drop_row[2] = {((static_cast<uint32_t>(b) * a.heads + h) * Sq + min(row[0], Sq - 1)) * Skv,
                                  ((static_cast<uint32_t>(b) * a.heads + h) * Sq + min(row[1], Sq - 1)) * Skv};
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const int a_off = (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (lane >> 4);

    for (int t0 = 0; t0 < w_end; t0 += MKT) {
        const bf16* k_tile = Ks + static_cast<size_t>(t0) * LD;
        const bf16* v_tile = Vs + static_cast<size_t>(t0) * LD;
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t qa[4], ga[4];
            ldsm_x4(qa, Qs + a_off + ks * 16);
            ldsm_x4(ga, Gs + a_off + ks * 16);
#pragma unroll
            for (int nb2 = 0; nb2 < 4; ++nb2) {
                const int b_off = (nb2 * 16 + (lane & 7) + 8 * (lane >> 4)) * LD + ks * 16 + 8 * ((lane >> 3) & 1);
                uint32_t kb[4], vb[4];
                ldsm_x4(kb, k_tile + b_off);
                ldsm_x4(vb, v_tile + b_off);
                mma_lp<FP16>(s[2 * nb2], qa, kb[0], kb[1]);
                mma_lp<FP16>(s[2 * nb2 + 1], qa, kb[2], kb[3]);
                mma_lp<FP16>(dp[2 * nb2], ga, vb[0], vb[1]);
                mma_lp<FP16>(dp[2 * nb2 + 1], ga, vb[2], vb[3]);
            }
        }
        // P and dS in place (s -> dS * ds_shift)
        const unsigned long long bits = s_bits[t0 >> 6];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                const int kc = nb * 8 + 2 * t + (e & 1), key = t0 + kc;
                const bool ok = (((key >= lo[r]) & (key < hi[r])) | (key == self[r])) & static_cast<bool>((bits >> kc) & 1ull);
                const float p = ok ? fast_exp2(fmaf(s[nb][e], sl, -lse2[r])) : 0.f;
                float dpv = dp[nb][e];
                if (a.drop.thresh != 0u)
                    dpv = drop_keep(a.drop.seed, drop_row[r] + static_cast<uint32_t>(key), a.drop.thresh) ? dpv * a.drop.scale : 0.f;
                s[nb][e] = p * (dpv - dl[r]) * a.ds_shift;
            }
        }
        // dQ += dS K
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t pa[4];
            pa[0] = pack2<FP16>(s[2 * kc][0], s[2 * kc][1]);
            pa[1] = pack2<FP16>(s[2 * kc][2], s[2 * kc][3]);
            pa[2] = pack2<FP16>(s[2 * kc + 1][0], s[2 * kc + 1][1]);
            pa[3] = pack2<FP16>(s[2 * kc + 1][2], s[2 * kc + 1][3]);
#pragma unroll
            for (int db2 = 0; db2 < D / 16; ++db2) {
                uint32_t kb[4];
                ldsm_x4_trans(kb, k_tile + (kc * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + db2 * 16 + 8 * (lane >> 4));
                mma_lp<FP16>(dq[2 * db2], pa, kb[0], kb[1]);
                mma_lp<FP16>(dq[2 * db2 + 1], pa, kb[2], kb[3]);
            }
        }
    }
    const float f = a.scale * a.inv_scale[0] / a.ds_shift;
    float* DQ = a.dq + static_cast<size_t>(b) * Sq * a.lddq + h * D;
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int col = i * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 4; ++e) { dq[i][e] *= f; mx = fmaxf(mx, fabsf(dq[i][e])); }
        if (row[0] < Sq) *reinterpret_cast<float2*>(DQ + static_cast<size_t>(row[0]) * a.lddq + col) = make_float2(dq[i][0], dq[i][1]);
        if (row[1] < Sq) *reinterpret_cast<float2*>(DQ + static_cast<size_t>(row[1]) * a.lddq + col) = make_float2(dq[i][2], dq[i][3]);
    }
    if (a.amax != nullptr) {       // rows beyond Sq hold zeros (their dO is zero-staged), so they cannot raise the maximum
        const unsigned u = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
        if (lane == 0) atomicMax(a.amax, u);
    }
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <int D, bool FP16, int NW>
__global__ void __launch_bounds__(NW * 32)
attn_bwd_dkv_kernel(BwdArgs a, int q_rows) {
    constexpr int MKT_CTA = 16 * NW, LD = D + PADE, NT = NW * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* Ks = reinterpret_cast<bf16*>(smem_raw);             // the CTA's key rows [MKT_CTA][LD]
    bf16* Vs = Ks + MKT_CTA * LD;
    bf16* Qs = Vs + MKT_CTA * LD;                             // every query row of the (sequence, head) [q_rows][LD]
    bf16* Gs = Qs + static_cast<size_t>(q_rows) * LD;         // dO, same
    float* s_lse = reinterpret_cast<float*>(Gs + static_cast<size_t>(q_rows) * LD);   // [q_rows] (log2 domain; +inf = padding)
    float* s_dl = s_lse + q_rows;                             // [q_rows]
    int* s_lo = reinterpret_cast<int*>(s_dl + q_rows);        // [q_rows] allowed key set of each query
    int* s_hi = s_lo + q_rows;
    int* s_self = s_hi + q_rows;
    __shared__ unsigned long long s_bits[4];

    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * MKT_CTA;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int Sq = a.Sq, Skv = a.Skv;
    const bf16* Q = a.q + static_cast<size_t>(b) * Sq * a.ldq + h * D;
    const bf16* K = a.k + static_cast<size_t>(b) * Skv * a.ldk + h * D;
    const bf16* V = a.v + static_cast<size_t>(b) * Skv * a.ldv + h * D;
    const bf16* G = a.dO + static_cast<size_t>(b) * Sq * a.lddo + h * D;

    SeqDesc desc = {0, 0, 0, 0};
    if (a.mask_kind != MASK_KEY_VECTOR) desc = a.desc[b];
    int q_real, k_real;
    real_extents(a.mask_kind, desc, Sq, Skv, q_real, k_real);
    if (k0 >= k_real) {                       // a block of keys no real query attends: zeros (block-uniform exit before any barrier)
        float* DK = a.dk + static_cast<size_t>(b) * Skv * a.lddk + h * D;
        float* DV = a.dv + static_cast<size_t>(b) * Skv * a.lddv + h * D;
        for (int i = tid; i < MKT_CTA * (D / 4); i += NT) {
            const int r = k0 + i / (D / 4), c = (i % (D / 4)) * 4;
            if (r < Skv) {
                *reinterpret_cast<float4*>(DK + static_cast<size_t>(r) * a.lddk + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(DV + static_cast<size_t>(r) * a.lddv + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        return;
    }
    const int q_lim = min(q_rows, ((q_real + MKT - 1) / MKT) * MKT);          // query tiles beyond hold padding rows only
    stage_rows<D>(Ks, K, a.ldk, k0, MKT_CTA, Skv, tid, NT);
    stage_rows<D>(Vs, V, a.ldv, k0, MKT_CTA, Skv, tid, NT);
    stage_rows<D>(Qs, Q, a.ldq, 0, q_lim, Sq, tid, NT);
    stage_rows<D>(Gs, G, a.lddo, 0, q_lim, Sq, tid, NT);
    const size_t stat0 = (static_cast<size_t>(b) * a.heads + h) * Sq;
    for (int r = tid; r < q_lim; r += NT) {
        int l = 0, hh = 0, sf = -1;
        if (r < Sq) row_set(a.mask_kind, desc, Skv, r, l, hh, sf);
        s_lo[r] = l; s_hi[r] = hh; s_self[r] = sf;
        s_lse[r] = r < Sq ? a.lse[stat0 + r] * kLog2e : INFINITY;
        s_dl[r] = r < Sq ? a.delta[stat0 + r] : 0.f;
    }
    // key bits of THIS CTA's key rows only are needed, but the helper builds all four words
    build_key_bits(a, b, ((Skv + 63) / 64) * 64, s_bits, tid, NT);
    cp_async_wait_all();
    __syncthreads();

    const int key[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
    bool key_ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) key_ok[r] = key[r] < Skv && ((s_bits[(key[r] >> 6) & 3] >> (key[r] & 63)) & 1ull);

    const float sl = a.scale * kLog2e;
    const uint32_t drop_bh = (static_cast<uint32_t>(b) * a.heads + h) * Sq;
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int a_off = (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (lane >> 4);

    for (int t0 = 0; t0 < q_lim; t0 += MKT) {
        const bf16* q_tile = Qs + static_cast<size_t>(t0) * LD;
        const bf16* g_tile = Gs + static_cast<size_t>(t0) * LD;
        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
        // S^T = K Q^T, dP^T = V dO^T  (rows = this warp's keys, columns = the tile's queries)
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t ka[4], va[4];
            ldsm_x4(ka, Ks + a_off + ks * 16);
            ldsm_x4(va, Vs + a_off + ks * 16);
#pragma unroll
            for (int nb2 = 0; nb2 < 4; ++nb2) {
                const int b_off = (nb2 * 16 + (lane & 7) + 8 * (lane >> 4)) * LD + ks * 16 + 8 * ((lane >> 3) & 1);
                uint32_t qb[4], gb[4];
                ldsm_x4(qb, q_tile + b_off);
                ldsm_x4(gb, g_tile + b_off);
                mma_lp<FP16>(s[2 * nb2], ka, qb[0], qb[1]);
                mma_lp<FP16>(s[2 * nb2 + 1], ka, qb[2], qb[3]);
                mma_lp<FP16>(dp[2 * nb2], va, gb[0], gb[1]);
                mma_lp<FP16>(dp[2 * nb2 + 1], va, gb[2], gb[3]);
            }
        }
        // P^T in s, dS^T * ds_shift in dp
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                const int qc = t0 + nb * 8 + 2 * t + (e & 1);
                const bool ok = key_ok[r] & (((key[r] >= s_lo[qc]) & (key[r] < s_hi[qc])) | (key[r] == s_self[qc]));
                const float p = ok ? fast_exp2(fmaf(s[nb][e], sl, -s_lse[qc])) : 0.f;
                float mk = 1.f;
                if (a.drop.thresh != 0u)
                    mk = drop_keep(a.drop.seed, (drop_bh + static_cast<uint32_t>(qc)) * Skv + static_cast<uint32_t>(key[r]), a.drop.thresh) ? a.drop.scale : 0.f;
                s[nb][e] = p * mk;                                       // P^T after dropout: the operand of dV
                dp[nb][e] = p * (dp[nb][e] * mk - s_dl[qc]) * a.ds_shift;
            }
        }
        // dV += P^T dO,  dK += dS^T Q   (contraction over the tile's 64 queries)
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t pa[4], da[4];
            pa[0] = pack2<FP16>(s[2 * kc][0], s[2 * kc][1]);
            pa[1] = pack2<FP16>(s[2 * kc][2], s[2 * kc][3]);
            pa[2] = pack2<FP16>(s[2 * kc + 1][0], s[2 * kc + 1][1]);
            pa[3] = pack2<FP16>(s[2 * kc + 1][2], s[2 * kc + 1][3]);
            da[0] = pack2<FP16>(dp[2 * kc][0], dp[2 * kc][1]);
            da[1] = pack2<FP16>(dp[2 * kc][2], dp[2 * kc][3]);
            da[2] = pack2<FP16>(dp[2 * kc + 1][0], dp[2 * kc + 1][1]);
            da[3] = pack2<FP16>(dp[2 * kc + 1][2], dp[2 * kc + 1][3]);
#pragma unroll
            for (int db2 = 0; db2 < D / 16; ++db2) {
                const int b_off = (kc * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + db2 * 16 + 8 * (lane >> 4);
                uint32_t gb[4], qb[4];
                ldsm_x4_trans(gb, g_tile + b_off);
                ldsm_x4_trans(qb, q_tile + b_off);
                mma_lp<FP16>(dv[2 * db2], pa, gb[0], gb[1]);
                mma_lp<FP16>(dv[2 * db2 + 1], pa, gb[2], gb[3]);
                mma_lp<FP16>(dk[2 * db2], da, qb[0], qb[1]);
                mma_lp<FP16>(dk[2 * db2 + 1], da, qb[2], qb[3]);
            }
        }
    }
    const float fv = a.inv_scale[0], fk = a.scale * a.inv_scale[0] / a.ds_shift;
    float* DK = a.dk + static_cast<size_t>(b) * Skv * a.lddk + h * D;
    float* DV = a.dv + static_cast<size_t>(b) * Skv * a.lddv + h * D;
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int col = i * 8 + 2 * t;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (key[r] < Skv) {
                const float2 k2 = make_float2(dk[i][2 * r] * fk, dk[i][2 * r + 1] * fk), v2 = make_float2(dv[i][2 * r] * fv, dv[i][2 * r + 1] * fv);
                *reinterpret_cast<float2*>(DK + static_cast<size_t>(key[r]) * a.lddk + col) = k2;
                *reinterpret_cast<float2*>(DV + static_cast<size_t>(key[r]) * a.lddv + col) = v2;
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(k2.x), fabsf(k2.y))), fmaxf(fabsf(v2.x), fabsf(v2.y)));
            }
        }
    }
    if (a.amax != nullptr) {
        const unsigned u = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));
        if (lane == 0) atomicMax(a.amax, u);
    }
}

template <int D, bool FP16, int NW>
int launch_dq(const BwdArgs& a, cudaStream_t stream) {
    constexpr int MQT = 16 * NW;
    const int kv_rows = ((a.Skv + MKT - 1) / MKT) * MKT;
    const size_t smem = sizeof(bf16) * (2 * MQT + 2 * static_cast<size_t>(kv_rows)) * (D + PADE);
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_bwd_dq_kernel<D, FP16, NW>), smem));
    dim3 grid((a.Sq + MQT - 1) / MQT, a.heads, a.B);
    attn_bwd_dq_kernel<D, FP16, NW><<<grid, NW * 32, smem, stream>>>(a, kv_rows);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

template <int D, bool FP16, int NW>
int launch_dkv(const BwdArgs& a, cudaStream_t stream) {
    constexpr int MKT_CTA = 16 * NW;
    const int q_rows = ((a.Sq + MKT - 1) / MKT) * MKT;
    const size_t smem = sizeof(bf16) * (2 * MKT_CTA + 2 * static_cast<size_t>(q_rows)) * (D + PADE) + static_cast<size_t>(q_rows) * 20;
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_bwd_dkv_kernel<D, FP16, NW>), smem));
    dim3 grid((a.Skv + MKT_CTA - 1) / MKT_CTA, a.heads, a.B);
    attn_bwd_dkv_kernel<D, FP16, NW><<<grid, NW * 32, smem, stream>>>(a, q_rows);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

template <int D, bool FP16>
int run_bwd(const BwdArgs& a, cudaStream_t stream) {
    // 128 rows per CTA when there are that many, 64 for the 37-region side
    if (a.Sq > 64) UNIMM_TRY((launch_dq<D, FP16, 8>(a, stream)));
    else UNIMM_TRY((launch_dq<D, FP16, 4>(a, stream)));
    // UNIMM_DKV_NW4=1 (A/B switch): 64 key rows per CTA everywhere -> two 4-warp CTAs per SM whose staging and compute phases overlap
    static const bool nw4 = getenv("UNIMM_DKV_NW4") != nullptr && atoi(getenv("UNIMM_DKV_NW4")) != 0;
    if (a.Skv > 64 && !nw4) UNIMM_TRY((launch_dkv<D, FP16, 8>(a, stream)));
    else UNIMM_TRY((launch_dkv<D, FP16, 4>(a, stream)));
    return 0;
}

}  // namespace

size_t attention_backward_scratch(int B, int heads, int D, int Sq) {
    const size_t rows = static_cast<size_t>(B) * Sq;
    return 256 + ((2 * rows * heads * D + 255) & ~size_t(255)) + ((4 * rows * heads + 255) & ~size_t(255));
}

int attention_backward_lp(const AttnArgs& f, const float* dO, int lddo, const float* lse, float* dq, int lddq, float* dk, int lddk,
                          float* dv, int lddv, void* scratch, size_t scratch_bytes, cudaStream_t stream, float* amax_accum, const float* dO_amax) {
    UNIMM_CHECK(f.B > 0 && f.B <= 65535 && f.heads > 0 && f.Sq > 0 && f.Skv > 0 && f.Sq <= 256 && f.Skv <= 256, "attention backward: bad problem size");
    UNIMM_CHECK(f.D == 64 || f.D == 128, "attention backward: head dim must be 64 or 128");
    UNIMM_CHECK((f.ldq % 8) == 0 && (f.ldk % 8) == 0 && (f.ldv % 8) == 0 && (f.ldo % 2) == 0, "attention backward: rows must be 16-byte aligned");
    UNIMM_CHECK((lddq % 4) == 0 && (lddk % 4) == 0 && (lddv % 4) == 0 && lddo == f.heads * f.D, "attention backward: dO must be contiguous [rows, heads * D], gradient rows 16-byte aligned");
    UNIMM_CHECK(((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0, "attention backward: 16-byte aligned gradient matrices");
    UNIMM_CHECK(f.mask_kind == MASK_KEY_VECTOR ? f.key_mask != nullptr : f.desc != nullptr, "attention backward: mask operand missing");
    UNIMM_CHECK(dO && lse && dq && dk && dv && scratch, "attention backward: null argument");
    UNIMM_CHECK(scratch_bytes >= attention_backward_scratch(f.B, f.heads, f.D, f.Sq), "attention backward: scratch too small");
    const size_t rows = static_cast<size_t>(f.B) * f.Sq;
    const int H = f.heads * f.D;
    char* p = static_cast<char*>(scratch);
    float* sc = reinterpret_cast<float*>(p);                                  // [0] = scale, [1] = 1 / scale
    bf16* dO16 = reinterpret_cast<bf16*>(p + 256);
    float* delta = reinterpret_cast<float*>(p + 256 + ((2 * rows * H + 255) & ~size_t(255)));
    UNIMM_TRY(amax_scale(dO, rows * H, f.lp_kind == LP_FP16 ? 1 : 0, sc, stream, dO_amax));
    UNIMM_TRY(cast_scaled_lp(dO, lddo, static_cast<int>(rows), H, sc, dO16, H, f.lp_kind, stream));
    {
        const size_t warps = rows * f.heads;
        attn_delta_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(dO16, H, static_cast<const bf16*>(f.o), f.ldo,
                                                                                             static_cast<int>(rows), f.Sq, f.heads, f.D, f.lp_kind, delta);
        UNIMM_LAUNCH_CHECK(1);
    }
    BwdArgs a;
    a.q = static_cast<const bf16*>(f.q); a.ldq = f.ldq; a.k = static_cast<const bf16*>(f.k); a.ldk = f.ldk;
    a.v = static_cast<const bf16*>(f.v); a.ldv = f.ldv; a.dO = dO16; a.lddo = H; a.lse = lse; a.delta = delta;
    a.dq = dq; a.lddq = lddq; a.dk = dk; a.lddk = lddk; a.dv = dv; a.lddv = lddv;
    a.B = f.B; a.heads = f.heads; a.Sq = f.Sq; a.Skv = f.Skv; a.mask_kind = f.mask_kind; a.desc = f.desc; a.key_mask = f.key_mask;
    a.scale = f.scale; a.inv_scale = sc + 1; a.ds_shift = f.lp_kind == LP_FP16 ? 0.0625f : 1.f;
    a.drop = f.drop;
    UNIMM_CHECK(f.drop.thresh == 0u || static_cast<double>(f.B) * f.heads * f.Sq * f.Skv < 4294967296.0, "attention dropout: 32-bit element index");
    a.amax = reinterpret_cast<unsigned*>(amax_accum);      // NOT zeroed here: several calls may fill column blocks of one gradient matrix
    if (f.lp_kind == LP_FP16) return f.D == 64 ? run_bwd<64, true>(a, stream) : run_bwd<128, true>(a, stream);
    return f.D == 64 ? run_bwd<64, false>(a, stream) : run_bwd<128, false>(a, stream);
}

}  // namespace unimm
