// Fused masked attention on the tensor cores (16-bit in, fp32 softmax/accumulate, 16-bit out).
//
// Same contract and mask generation as attention.cu (the CUDA-core kernel it is tested against); this
// is the throughput version used by the bf16 / fp16 engine modes.  One CTA = NW warps = 16*NW query rows
// of one (sequence, head); each warp owns 16 rows.  The keys/values the CTA's rows may attend (union of
// their descriptor intervals, rounded to 64) are staged once in shared memory with cp.async; every warp
// then runs a FlashAttention-2 style loop over the 64-key tiles ITS rows need: S = Q K^T with
// mma.sync.m16n8k16, on-the-fly mask, online softmax in registers (scale folded into the exp2 FFMA),
// P re-used as the A fragment of O += P V.
//
// The mask never exists as a tensor: per 64-key tile each thread turns its two rows' descriptor
// intervals into two 64-bit masks (a few integer ops), tiles that are fully allowed for the whole warp
// skip masking altogether, and tiles beyond a warp's last allowed key are never visited.
#include "attn_common.cuh"
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

using namespace attn;

// Allowed set of query row qr.  Padding rows (no allowed key in the reference) attend key 0 only: their
// output is never read by a valid row, it only has to be finite and cheap.
__device__ __forceinline__ void row_set(const AttnArgs& a, const SeqDesc& desc, int qr, int& lo, int& hi, int& self) {
    self = -1;
    if (a.mask_kind == MASK_TEXT_SELF) {
        text_row_interval(desc, qr, a.Skv, lo, hi, self);
        if (hi <= lo && self < 0) { lo = 0; hi = 1; }
    } else if (a.mask_kind == MASK_CO_INTERVAL) {
        co_interval(desc, a.Skv, lo, hi);
        if (hi <= lo) { lo = 0; hi = a.Skv; }       // reference: every column shifted by -10000 -> plain softmax
    } else {
        lo = 0; hi = a.Skv;
    }
}

template <int D, bool FP16, int NW>
__global__ void __launch_bounds__(NW * 32)
attn_mma_kernel(AttnArgs a, int kv_rows_max) {
    constexpr int MQT = 16 * NW;       // query rows per CTA
    constexpr int LD = D + PADE;       // smem row stride in elements
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);          // [MQT][LD]
    bf16* Ks = Qs + MQT * LD;                              // [kv_rows_max][LD]
    bf16* Vs = Ks + static_cast<size_t>(kv_rows_max) * LD; // [kv_rows_max][LD]
    __shared__ unsigned long long s_keymask[4];            // KEY_VECTOR: allowed keys per 64-key tile (Skv <= 256)
    __shared__ int s_kv_hi;

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MQT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int Sq = a.Sq, Skv = a.Skv;
    const bf16* Q = static_cast<const bf16*>(a.q) + static_cast<size_t>(b) * Sq * a.ldq + h * D;
    const bf16* K = static_cast<const bf16*>(a.k) + static_cast<size_t>(b) * Skv * a.ldk + h * D;
    const bf16* V = static_cast<const bf16*>(a.v) + static_cast<size_t>(b) * Skv * a.ldv + h * D;
    bf16* O = static_cast<bf16*>(a.o) + static_cast<size_t>(b) * Sq * a.ldo + h * D;

    // ---- allowed key sets: lanes 0..15 each evaluate one of the warp's rows, the warp reduces its key range
    SeqDesc desc = {0, 0, 0, 0};
    if (a.mask_kind != MASK_KEY_VECTOR) desc = a.desc[b];
    if (tid == 0) s_kv_hi = 0;
    if (tid < 4) s_keymask[tid] = 0ull;
    __syncthreads();
    int w_hi = 0;
    {
        const int qr = q0 + warp * 16 + (lane & 15);
        if (qr < Sq) {
            int l, hh, s;
            row_set(a, desc, qr, l, hh, s);
            w_hi = max(hh, s + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
        if (lane == 0 && w_hi > 0) atomicMax(&s_kv_hi, w_hi);
    }
    int lo[2], hi[2], self[2];
    row_set(a, desc, min(q0 + warp * 16 + g, Sq - 1), lo[0], hi[0], self[0]);
    row_set(a, desc, min(q0 + warp * 16 + g + 8, Sq - 1), lo[1], hi[1], self[1]);
    if (a.mask_kind == MASK_KEY_VECTOR) {
        // one ballot per 32 keys builds the per-tile allowed-key masks; all-masked rows fall back to "all keys"
        const float* km = a.key_mask + static_cast<size_t>(b) * Skv;
        for (int k0 = warp * 32; k0 < kv_rows_max; k0 += NW * 32) {
            const int key = k0 + lane;
            const unsigned bits = __ballot_sync(0xffffffffu, key < Skv && km[key] > 0.5f);
            if (lane == 0 && bits) atomicOr(&s_keymask[k0 >> 6], static_cast<unsigned long long>(bits) << (k0 & 32));
        }
    }
    __syncthreads();
    const int kv_hi = s_kv_hi;                               // CTA-wide: rows [0, roundup64(kv_hi)) are staged
    const int n_rows = min(((kv_hi + MKT - 1) / MKT) * MKT, kv_rows_max);
    const int w_end = min(((w_hi + MKT - 1) / MKT) * MKT, n_rows);   // this warp's last tile end

    // ---- stage the Q tile and K/V rows [0, n_rows) (rows >= Skv are zero filled)
    constexpr int CH = D / 8;   // 16-byte chunks per row
    for (int i = tid; i < MQT * CH; i += NT) {
        const int r = i / CH, c = (i % CH) * 8;
        bf16* dst = Qs + r * LD + c;
        if (q0 + r < Sq) cp_async16(dst, Q + static_cast<size_t>(q0 + r) * a.ldq + c);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < n_rows * CH; i += NT) {
        const int r = i / CH, c = (i % CH) * 8;
        bf16* dk = Ks + r * LD + c;
        bf16* dv = Vs + r * LD + c;
        if (r < Skv) {
            cp_async16(dk, K + static_cast<size_t>(r) * a.ldk + c);
            cp_async16(dv, V + static_cast<size_t>(r) * a.ldv + c);
        } else {
            *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    bool key_all = false;
    if (a.mask_kind == MASK_KEY_VECTOR)
        key_all = (s_keymask[0] | s_keymask[1] | s_keymask[2] | s_keymask[3]) == 0ull;   // every key masked: uniform shift only

    // ---- main loop over this warp's tiles
    const float sl = a.scale * 1.4426950408889634f;          // softmax scale folded into the exp2 argument
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

    const bf16* q_base = Qs + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (lane >> 4);
    for (int t0 = 0; t0 < w_end; t0 += MKT) {
        const bf16* k_tile = Ks + static_cast<size_t>(t0) * LD;
        const bf16* v_tile = Vs + static_cast<size_t>(t0) * LD;
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t qa[4];
            ldsm_x4(qa, q_base + ks * 16);
#pragma unroll
            for (int nb2 = 0; nb2 < 4; ++nb2) {
                uint32_t kb[4];
                ldsm_x4(kb, k_tile + (nb2 * 16 + (lane & 7) + 8 * (lane >> 4)) * LD + ks * 16 + 8 * ((lane >> 3) & 1));
                mma_lp<FP16>(s[2 * nb2], qa, kb[0], kb[1]);
                mma_lp<FP16>(s[2 * nb2 + 1], qa, kb[2], kb[3]);
            }
        }
        // ---- mask: two 64-bit masks per thread (its rows g and g+8); skipped when the warp's tile is fully allowed
        unsigned long long m0, m1;
        if (a.mask_kind == MASK_KEY_VECTOR) {
            m0 = m1 = key_all ? tile_mask(0, Skv, -1, t0) : s_keymask[t0 >> 6];
        } else {
            m0 = tile_mask(lo[0], min(hi[0], Skv), self[0], t0);
            m1 = tile_mask(lo[1], min(hi[1], Skv), self[1], t0);
        }
        if (!__all_sync(0xffffffffu, (m0 & m1) == ~0ull)) {
            m0 >>= 2 * t;
            m1 >>= 2 * t;
            const uint32_t a0 = static_cast<uint32_t>(m0), a1 = static_cast<uint32_t>(m0 >> 32);
            const uint32_t b0 = static_cast<uint32_t>(m1), b1 = static_cast<uint32_t>(m1 >> 32);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const uint32_t wa = nb < 4 ? a0 : a1, wb = nb < 4 ? b0 : b1;
                const int sh = (nb & 3) * 8;
                if (!((wa >> sh) & 1u)) s[nb][0] = -INFINITY;
                if (!((wa >> (sh + 1)) & 1u)) s[nb][1] = -INFINITY;
                if (!((wb >> sh) & 1u)) s[nb][2] = -INFINITY;
                if (!((wb >> (sh + 1)) & 1u)) s[nb][3] = -INFINITY;
            }
        }
        // ---- online softmax on raw scores; p = exp2(s*sl - m*sl)
        float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            tmax[0] = fmaxf(tmax[0], fmaxf(s[nb][0], s[nb][1]));
            tmax[1] = fmaxf(tmax[1], fmaxf(s[nb][2], s[nb][3]));
        }
        float corr[2], msl[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
            const float m_new = fmaxf(m_run[r], tmax[r]);
            corr[r] = (m_new == -INFINITY) ? 1.f : fast_exp2((m_run[r] - m_new) * sl);
            m_run[r] = m_new;
            msl[r] = (m_new == -INFINITY) ? 0.f : m_new * sl;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = fast_exp2(fmaf(s[nb][0], sl, -msl[0]));   // masked: fma(-inf, sl, x) = -inf -> 0
            s[nb][1] = fast_exp2(fmaf(s[nb][1], sl, -msl[0]));
            s[nb][2] = fast_exp2(fmaf(s[nb][2], sl, -msl[1]));
            s[nb][3] = fast_exp2(fmaf(s[nb][3], sl, -msl[1]));
            l_run[0] += s[nb][0] + s[nb][1];
            l_run[1] += s[nb][2] + s[nb][3];
        }
        if (a.drop.thresh != 0u) {          // training: dropout on the probabilities (the row sums above stay those of the full softmax)
            const uint32_t r0i = ((static_cast<uint32_t>(b) * a.heads + h) * Sq + min(q0 + warp * 16 + g, Sq - 1)) * Skv + t0 + 2 * t;
            const uint32_t r1i = ((static_cast<uint32_t>(b) * a.heads + h) * Sq + min(q0 + warp * 16 + g + 8, Sq - 1)) * Skv + t0 + 2 * t;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                if (!drop_keep(a.drop.seed, r0i + nb * 8, a.drop.thresh)) s[nb][0] = 0.f;
                if (!drop_keep(a.drop.seed, r0i + nb * 8 + 1, a.drop.thresh)) s[nb][1] = 0.f;
                if (!drop_keep(a.drop.seed, r1i + nb * 8, a.drop.thresh)) s[nb][2] = 0.f;
                if (!drop_keep(a.drop.seed, r1i + nb * 8 + 1, a.drop.thresh)) s[nb][3] = 0.f;
            }
        }
        if (__any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f)) {
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                o[i][0] *= corr[0]; o[i][1] *= corr[0];
                o[i][2] *= corr[1]; o[i][3] *= corr[1];
            }
        }
        // ---- O += P V
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t pa[4];
            pa[0] = pack2<FP16>(s[2 * kc][0], s[2 * kc][1]);
            pa[1] = pack2<FP16>(s[2 * kc][2], s[2 * kc][3]);
            pa[2] = pack2<FP16>(s[2 * kc + 1][0], s[2 * kc + 1][1]);
            pa[3] = pack2<FP16>(s[2 * kc + 1][2], s[2 * kc + 1][3]);
#pragma unroll
            for (int db2 = 0; db2 < D / 16; ++db2) {
                uint32_t vb[4];
                ldsm_x4_trans(vb, v_tile + (kc * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + db2 * 16 + 8 * (lane >> 4));
                mma_lp<FP16>(o[2 * db2], pa, vb[0], vb[1]);
                mma_lp<FP16>(o[2 * db2 + 1], pa, vb[2], vb[3]);
            }
        }
    }
    // ---- finalize: row sums across the quad, normalise, store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? a.drop.scale / l_run[0] : 0.f;      // drop.scale = 1 / (1 - p), 1 without dropout
    const float inv1 = l_run[1] > 0.f ? a.drop.scale / l_run[1] : 0.f;
    const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int col = i * 8 + 2 * t;
        if (row0 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row0) * a.ldo + col) = pack2<FP16>(o[i][0] * inv0, o[i][1] * inv0);
        if (row1 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row1) * a.ldo + col) = pack2<FP16>(o[i][2] * inv1, o[i][3] * inv1);
    }
    if (a.lse != nullptr && t == 0) {   // natural-log domain, softmax scale included: what the backward recomputes P from
        float* L = a.lse + (static_cast<size_t>(b) * a.heads + h) * Sq;
        if (row0 < Sq) L[row0] = m_run[0] * a.scale + logf(l_run[0]);
        if (row1 < Sq) L[row1] = m_run[1] * a.scale + logf(l_run[1]);
    }
}

template <int D, bool FP16, int NW>
int launch_mma(const AttnArgs& a, cudaStream_t stream) {
    constexpr int MQT = 16 * NW;
    const int kv_rows_max = ((a.Skv + MKT - 1) / MKT) * MKT;
    const size_t smem = sizeof(bf16) * (MQT + 2 * static_cast<size_t>(kv_rows_max)) * (D + PADE);
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_mma_kernel<D, FP16, NW>), smem));
    dim3 grid((a.Sq + MQT - 1) / MQT, a.heads, a.B);
    attn_mma_kernel<D, FP16, NW><<<grid, NW * 32, smem, stream>>>(a, kv_rows_max);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

template <int D, bool FP16>
int dispatch_rows(const AttnArgs& a, cudaStream_t stream) {
    // 128 query rows per CTA share one staged K/V when there are that many; short query sets (37 regions) use 64
    return a.Sq > 64 ? launch_mma<D, FP16, 8>(a, stream) : launch_mma<D, FP16, 4>(a, stream);
}

}  // namespace

int attention_mma_lp(const AttnArgs& a, cudaStream_t stream) {
    UNIMM_CHECK(a.B > 0 && a.B <= 65535 && a.heads > 0 && a.Sq > 0 && a.Skv > 0 && a.Skv <= 256, "attention: bad problem size");
    UNIMM_CHECK(a.D == 64 || a.D == 128, "attention: head dim must be 64 or 128");
    UNIMM_CHECK((a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 2) == 0, "attention: rows must be 16-byte aligned");
    UNIMM_CHECK(a.mask_kind == MASK_KEY_VECTOR ? a.key_mask != nullptr : a.desc != nullptr, "attention: mask operand missing");
    if (a.lp_kind == LP_FP16) return a.D == 64 ? dispatch_rows<64, true>(a, stream) : dispatch_rows<128, true>(a, stream);
    return a.D == 64 ? dispatch_rows<64, false>(a, stream) : dispatch_rows<128, false>(a, stream);
}

}  // namespace unimm
