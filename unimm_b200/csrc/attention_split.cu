// fp32-class attention on the tensor cores (the "fp32" precision mode): the job semantics of attention_jobs.cu with every
// operand as TWO fp16 planes (x = hi + lo, LP_HILO: 22 mantissa bits) and three mma.sync passes per product,
//     S = Q_lo K_hi^T + Q_hi K_lo^T + Q_hi K_hi^T        O = P_lo V_hi + P_hi V_lo + P_hi V_hi,
// accumulated in fp32 registers; the softmax itself is fp32 (scale before mask, reference models/vilbert_dialog.py:395-410,
// :524-539, :681-721).  Q / K / V planes are what the split3 QKV GEMM wrote (gemm_umma.cu, out_hilo); the context goes out as
// planes too, so the output projection reads it without a conversion pass.  Replaces attn_jobs_simt_kernel (one warp per query
// row on the CUDA cores: 76 % of the fp32-mode step) — same masks, same "no valid key = all keys" rule.
//
// One CTA = NW warps = 16 NW query rows of one (job, head).  Keys are streamed in 64-row chunks (both planes of K and V:
// 37 KB at D = 64), so no capacity limit applies to the key range or the window; 2-3 CTAs per SM overlap each other's loads.
#include "attn_common.cuh"
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

using namespace attn;

template <int D, int NW>
__global__ void __launch_bounds__(NW * 32, 2)
attn_jobs_split_kernel(AttnJobsArgs a) {
    constexpr int MQT = 16 * NW;
    constexpr int LD = D + PADE;
    constexpr int NT = NW * 32;
    constexpr int CH = D / 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* Qh = reinterpret_cast<bf16*>(smem_raw);           // [MQT][LD]
    bf16* Ql = Qh + MQT * LD;
    bf16* Kh = Ql + MQT * LD;                               // [MKT][LD] each
    bf16* Kl = Kh + MKT * LD;
    bf16* Vh = Kl + MKT * LD;
    bf16* Vl = Vh + MKT * LD;
    __shared__ unsigned long long s_keymask[4];
    __shared__ int s_wlo, s_whi;

    const int* job = a.jobs + static_cast<size_t>(blockIdx.z) * 8;
    const int q_start = job[0], q_len = job[1], kv_start = job[2], kv_len = job[3], win = job[4], mask_row = job[5];
    const int h = blockIdx.y, q0 = blockIdx.x * MQT;
    if (q0 >= q_len) return;                                // block-uniform
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const bf16* Q = static_cast<const bf16*>(a.q) + static_cast<size_t>(q_start) * a.ldq + h * D;
    const bf16* K = static_cast<const bf16*>(a.k) + h * D;
    const bf16* V = static_cast<const bf16*>(a.v) + h * D;
    bf16* O = static_cast<bf16*>(a.o) + static_cast<size_t>(q_start) * a.ldo + h * D;
    const int n1p = ((kv_len + MKT - 1) / MKT) * MKT;       // range 1 padded to whole tiles

    if (tid == 0) { s_wlo = 0x7fffffff; s_whi = 0; }
    if (tid < 4) s_keymask[tid] = 0ull;
    __syncthreads();
    int lo[2] = {0, 0}, hi[2] = {0, 0}, self[2] = {-1, -1};
    int w_lo = 0x7fffffff, w_hi = 0;
    if (win) {
        const int qr = q0 + warp * 16 + (lane & 15);
        if (qr < q_len) {
            const int4 iv = *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(q_start + qr) * 4);
            w_lo = min(iv.x, iv.z >= 0 ? iv.z : iv.x);
            w_hi = max(iv.y, iv.z + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
            w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
        }
        if (lane == 0 && w_hi > w_lo) { atomicMin(&s_wlo, w_lo); atomicMax(&s_whi, w_hi); }
        const int r0 = min(q0 + warp * 16 + g, q_len - 1), r1 = min(q0 + warp * 16 + g + 8, q_len - 1);
        const int4 i0 = *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(q_start + r0) * 4);
        const int4 i1 = *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(q_start + r1) * 4);
        lo[0] = i0.x; hi[0] = i0.y; self[0] = i0.z;
        lo[1] = i1.x; hi[1] = i1.y; self[1] = i1.z;
    }
    if (mask_row >= 0) {
        const float* km = a.key_mask + static_cast<size_t>(mask_row) * a.key_mask_ld;
        for (int k0 = warp * 32; k0 < n1p; k0 += NW * 32) {
            const int key = k0 + lane;
            const unsigned bits = __ballot_sync(0xffffffffu, key < kv_len && km[key] > 0.5f);
            if (lane == 0 && bits) atomicOr(&s_keymask[k0 >> 6], static_cast<unsigned long long>(bits) << (k0 & 32));
        }
    }
    // ---- Q planes (once)
    for (int i = tid; i < MQT * CH; i += NT) {
        const int r = i / CH, c = (i % CH) * 8;
        if (q0 + r < q_len) {
            const bf16* src = Q + static_cast<size_t>(q0 + r) * a.ldq + c;
            cp_async16(Qh + r * LD + c, src);
            cp_async16(Ql + r * LD + c, src + a.lo_off_q);
        } else {
            *reinterpret_cast<uint4*>(Qh + r * LD + c) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(Ql + r * LD + c) = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    const int c_wlo = s_wlo, c_whi = s_whi;                  // CTA window [c_wlo, c_whi) in packed rows
    const int n2 = (win && c_whi > c_wlo) ? (c_whi - c_wlo) : 0;
    const int n2p = ((n2 + MKT - 1) / MKT) * MKT;
    const int n_rows = n1p + n2p;
    bool key_all = false;
    if (mask_row >= 0) key_all = (s_keymask[0] | s_keymask[1] | s_keymask[2] | s_keymask[3]) == 0ull;

    const int shift = n1p - c_wlo;
    int blo[2], bhi[2], bself[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        blo[r] = win ? lo[r] + shift : 0;
        bhi[r] = win ? hi[r] + shift : 0;
        bself[r] = (win && self[r] >= 0) ? self[r] + shift : -1;
    }
    // staged rows this WARP needs: range 1, and the window up to its own rows' last key
    const bool has_w = win && w_hi > w_lo;
    const int w_end = has_w ? min(((max(w_hi + shift, kv_len) + MKT - 1) / MKT) * MKT, n_rows) : n1p;
    const int w_begin2 = has_w ? max(n1p, ((w_lo + shift) / MKT) * MKT) : n_rows;     // first window tile this warp needs

    const float sl = a.scale * 1.4426950408889634f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

    const int q_off = (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (lane >> 4);
    for (int t0 = 0; t0 < n_rows; t0 += MKT) {
        __syncthreads();                                     // the previous chunk has been consumed by every warp
        for (int i = tid; i < MKT * CH; i += NT) {
            const int r = i / CH, c = (i % CH) * 8, row = t0 + r;
            int src = -1;
            if (row < kv_len) src = kv_start + row;
            else if (row >= n1p && row - n1p < n2) src = c_wlo + (row - n1p);
            const int d = r * LD + c;
            if (src >= 0) {
                const bf16* ks = K + static_cast<size_t>(src) * a.ldk + c;
                const bf16* vs = V + static_cast<size_t>(src) * a.ldv + c;
                cp_async16(Kh + d, ks); cp_async16(Kl + d, ks + a.lo_off_k);
                cp_async16(Vh + d, vs); cp_async16(Vl + d, vs + a.lo_off_v);
            } else {
                const uint4 z = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(Kh + d) = z; *reinterpret_cast<uint4*>(Kl + d) = z;
                *reinterpret_cast<uint4*>(Vh + d) = z; *reinterpret_cast<uint4*>(Vl + d) = z;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (t0 >= w_end || (t0 >= n1p && t0 < w_begin2)) continue;      // warp-uniform: none of this warp's rows has a key here

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t qh[4], ql[4];
            ldsm_x4(qh, Qh + q_off + ks * 16);
            ldsm_x4(ql, Ql + q_off + ks * 16);
#pragma unroll
            for (int nb2 = 0; nb2 < 4; ++nb2) {
                const int k_off = (nb2 * 16 + (lane & 7) + 8 * (lane >> 4)) * LD + ks * 16 + 8 * ((lane >> 3) & 1);
                uint32_t kh[4], kl[4];
                ldsm_x4(kh, Kh + k_off);
                ldsm_x4(kl, Kl + k_off);
                mma_lp<true>(s[2 * nb2], ql, kh[0], kh[1]);          // small terms first
                mma_lp<true>(s[2 * nb2], qh, kl[0], kl[1]);
                mma_lp<true>(s[2 * nb2], qh, kh[0], kh[1]);
                mma_lp<true>(s[2 * nb2 + 1], ql, kh[2], kh[3]);
                mma_lp<true>(s[2 * nb2 + 1], qh, kl[2], kl[3]);
                mma_lp<true>(s[2 * nb2 + 1], qh, kh[2], kh[3]);
            }
        }
        unsigned long long m0, m1;
        if (mask_row >= 0 && !key_all && t0 < n1p) {
            m0 = m1 = s_keymask[t0 >> 6];
        } else {
            m0 = tile_mask2(0, kv_len, blo[0], bhi[0], bself[0], t0);
            m1 = tile_mask2(0, kv_len, blo[1], bhi[1], bself[1], t0);
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
            corr[r] = (m_new == -INFINITY) ? 1.f : exp2f((m_run[r] - m_new) * sl);
            m_run[r] = m_new;
            msl[r] = (m_new == -INFINITY) ? 0.f : m_new * sl;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = exp2f(fmaf(s[nb][0], sl, -msl[0]));
            s[nb][1] = exp2f(fmaf(s[nb][1], sl, -msl[0]));
            s[nb][2] = exp2f(fmaf(s[nb][2], sl, -msl[1]));
            s[nb][3] = exp2f(fmaf(s[nb][3], sl, -msl[1]));
            l_run[0] += s[nb][0] + s[nb][1];
            l_run[1] += s[nb][2] + s[nb][3];
        }
        if (__any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f)) {
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                o[i][0] *= corr[0]; o[i][1] *= corr[0];
                o[i][2] *= corr[1]; o[i][3] *= corr[1];
            }
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t ph[4], pl[4];
            split_hilo2(s[2 * kc][0], s[2 * kc][1], ph[0], pl[0]);
            split_hilo2(s[2 * kc][2], s[2 * kc][3], ph[1], pl[1]);
            split_hilo2(s[2 * kc + 1][0], s[2 * kc + 1][1], ph[2], pl[2]);
            split_hilo2(s[2 * kc + 1][2], s[2 * kc + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int db2 = 0; db2 < D / 16; ++db2) {
                const int v_off = (kc * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + db2 * 16 + 8 * (lane >> 4);
                uint32_t vh[4], vl[4];
                ldsm_x4_trans(vh, Vh + v_off);
                ldsm_x4_trans(vl, Vl + v_off);
                mma_lp<true>(o[2 * db2], pl, vh[0], vh[1]);
                mma_lp<true>(o[2 * db2], ph, vl[0], vl[1]);
                mma_lp<true>(o[2 * db2], ph, vh[0], vh[1]);
                mma_lp<true>(o[2 * db2 + 1], pl, vh[2], vh[3]);
                mma_lp<true>(o[2 * db2 + 1], ph, vl[2], vl[3]);
                mma_lp<true>(o[2 * db2 + 1], ph, vh[2], vh[3]);
            }
        }
    }
    cp_async_wait_all();                                     // (a job without any key never entered the loop)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int col = i * 8 + 2 * t;
        uint32_t vh, vl;
        if (row0 < q_len) {
            split_hilo2(o[i][0] * inv0, o[i][1] * inv0, vh, vl);
            bf16* dst = O + static_cast<size_t>(row0) * a.ldo + col;
            *reinterpret_cast<uint32_t*>(dst) = vh;
            *reinterpret_cast<uint32_t*>(dst + a.lo_off_o) = vl;
        }
        if (row1 < q_len) {
            split_hilo2(o[i][2] * inv1, o[i][3] * inv1, vh, vl);
            bf16* dst = O + static_cast<size_t>(row1) * a.ldo + col;
            *reinterpret_cast<uint32_t*>(dst) = vh;
            *reinterpret_cast<uint32_t*>(dst + a.lo_off_o) = vl;
        }
    }
}

template <int D, int NW>
int launch_split(const AttnJobsArgs& a, cudaStream_t stream) {
    constexpr int MQT = 16 * NW;
    const size_t smem = sizeof(bf16) * (2 * MQT + 4 * MKT) * (D + PADE);
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_jobs_split_kernel<D, NW>), smem));
    dim3 grid((a.max_q_len + MQT - 1) / MQT, a.heads, a.n_jobs);
    attn_jobs_split_kernel<D, NW><<<grid, NW * 32, smem, stream>>>(a);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

// dense [B, S] layout -> the job lists / per-row intervals the kernel above consumes (the masks of utils/data_utils.py:149-210,
// :300-354 from the 4-integer descriptors): text self-attention as window-only jobs (row r of sequence b sees the rows
// b S + [lo, hi) U {self}; a padding row sees the whole sequence, as the reference's additive -10000 leaves it), and the
// image -> text co-attention interval as the shared key range of (b R, R) queries.
__global__ void dense_jobs_kernel(const SeqDesc* __restrict__ desc, int B, int S, int R, int* __restrict__ jobs_text,
                                  int* __restrict__ jobs_i2t, int* __restrict__ jobs_img, int* __restrict__ row_iv) {
    const int b = blockIdx.x;
    const SeqDesc d = desc[b];
    for (int r = threadIdx.x; r < S; r += blockDim.x) {
        int lo, hi, self;
        text_row_interval(d, r, S, lo, hi, self);
        if (hi <= lo && self < 0) { lo = 0; hi = S; }
        int* iv = row_iv + (static_cast<size_t>(b) * S + r) * 4;
        iv[0] = b * S + lo; iv[1] = b * S + hi; iv[2] = self >= 0 ? b * S + self : -1; iv[3] = 0;
    }
    if (threadIdx.x == 0) {
        int* j = jobs_text + b * 8;
        j[0] = b * S; j[1] = S; j[2] = b * S; j[3] = 0; j[4] = 1; j[5] = -1; j[6] = 0; j[7] = 0;
        int lo, hi;
        co_interval(d, S, lo, hi);
        if (hi <= lo) { lo = 0; hi = S; }
        j = jobs_i2t + b * 8;
        j[0] = b * R; j[1] = R; j[2] = b * S + lo; j[3] = hi - lo; j[4] = 0; j[5] = -1; j[6] = 0; j[7] = 0;
        j = jobs_img + b * 8;
        j[0] = b * R; j[1] = R; j[2] = b * R; j[3] = R; j[4] = 0; j[5] = b; j[6] = 0; j[7] = 0;
    }
}

}  // namespace

int attention_jobs_split(const AttnJobsArgs& a, cudaStream_t stream) {
    UNIMM_CHECK(a.n_jobs > 0 && a.n_jobs <= 65535 && a.heads > 0 && a.max_q_len > 0, "split attention: bad problem size");
    UNIMM_CHECK(a.D == 64 || a.D == 128, "split attention: head dim must be 64 or 128");
    UNIMM_CHECK((a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 2) == 0 && (a.lo_off_q % 8) == 0 &&
                    (a.lo_off_k % 8) == 0 && (a.lo_off_v % 8) == 0 && (a.lo_off_o % 2) == 0 && a.lo_off_q > 0 && a.lo_off_k > 0 &&
                    a.lo_off_v > 0 && a.lo_off_o > 0, "split attention: planes must be 16-byte aligned");
    if (a.D == 64) return a.max_q_len > 64 ? launch_split<64, 8>(a, stream) : launch_split<64, 4>(a, stream);
    return launch_split<128, 4>(a, stream);
}

int build_dense_jobs(const SeqDesc* desc, int B, int S, int R, int* jobs_text, int* jobs_i2t, int* jobs_img, int* row_iv,
                     cudaStream_t stream) {
    dense_jobs_kernel<<<B, 128, 0, stream>>>(desc, B, S, R, jobs_text, jobs_i2t, jobs_img, row_iv);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace unimm
