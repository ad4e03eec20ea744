// Job-based fused attention for the PREFIX-SHARED generative-scoring layout (tensor cores, 16-bit in/out).
//
// In generative mode the context rows [1,ctx) and all image rows are bit-identical across the 100
// candidates of a dialog round (context rows attend only context + image, image rows only
// context).  The packed layout therefore stores, per (image, round) unit, ONE copy of the context rows and,
// per candidate, only its own rows [CLS, A_0..A_{last-1}, B_0..B_{last-1}] (utils/data_utils.py:199-210).
// Every attention of the encoder then is a list of JOBS over packed rows:
//
//   (q_start, q_len)    query rows            (kv_start, kv_len)   key/value rows, all allowed
//   win = 1             additionally each query row r may attend its own candidate's rows: an interval
//                       [lo_r, hi_r) of packed rows plus optionally itself (row_iv[r] = lo, hi, self) —
//                       CLS: all own rows; A_k: A_0..A_k; B_k: A_0..A_{k-1} and itself
//   mask_row >= 0       key validity from key_mask[mask_row] (image padding mask of the unit)
//
// One CTA = NW warps = 16*NW query rows of one (job, head).  It stages the job's K/V range once, then the
// window [min lo_r, max hi_r) of its rows behind it (64-aligned), and runs the same FlashAttention-2 loop
// as attention_mma.cu with 64-bit tile masks built from  [0,kv_len) U [lo_r,hi_r) U {self}.
#include "attn_common.cuh"
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

using namespace attn;

template <int D, bool FP16, int NW>
__global__ void __launch_bounds__(NW * 32)
attn_jobs_kernel(AttnJobsArgs a) {
    constexpr int MQT = 16 * NW;
    constexpr int LD = D + PADE;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int rows_cap = a.kv_cap + a.win_cap;
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);           // [MQT][LD]
    bf16* Ks = Qs + MQT * LD;                               // [rows_cap][LD]
    bf16* Vs = Ks + static_cast<size_t>(rows_cap) * LD;     // [rows_cap][LD]
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

    // ---- window of the CTA's rows (packed row indices), reduced over lanes then warps
    if (tid == 0) { s_wlo = 0x7fffffff; s_whi = 0; }
    if (tid < 4) s_keymask[tid] = 0ull;
    __syncthreads();
    int lo[2] = {0, 0}, hi[2] = {0, 0}, self[2] = {-1, -1};   // this thread's two rows, packed-row coordinates
    int w_lo = 0x7fffffff, w_hi = 0;                         // this warp's window
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
    __syncthreads();
    const int c_wlo = s_wlo, c_whi = s_whi;                  // CTA window [c_wlo, c_whi) in packed rows
    const int n2 = (win && c_whi > c_wlo) ? min(c_whi - c_wlo, a.win_cap) : 0;
    const int n2p = ((n2 + MKT - 1) / MKT) * MKT;
    const int n_rows = n1p + n2p;

    // ---- stage Q, the shared K/V range, then the window rows
    constexpr int CH = D / 8;
    for (int i = tid; i < MQT * CH; i += NT) {
        const int r = i / CH, c = (i % CH) * 8;
        bf16* dst = Qs + r * LD + c;
        if (q0 + r < q_len) cp_async16(dst, Q + static_cast<size_t>(q0 + r) * a.ldq + c);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < n_rows * CH; i += NT) {
        const int r = i / CH, c = (i % CH) * 8;
        int src = -1;
        if (r < kv_len) src = kv_start + r;
        else if (r >= n1p && r - n1p < n2) src = c_wlo + (r - n1p);
        bf16* dk = Ks + r * LD + c;
        bf16* dv = Vs + r * LD + c;
        if (src >= 0) {
            cp_async16(dk, K + static_cast<size_t>(src) * a.ldk + c);
            cp_async16(dv, V + static_cast<size_t>(src) * a.ldv + c);
        } else {
            *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    bool key_all = false;
    if (mask_row >= 0) key_all = (s_keymask[0] | s_keymask[1] | s_keymask[2] | s_keymask[3]) == 0ull;

    // staged-buffer coordinates of the per-row window intervals
    const int shift = n1p - c_wlo;
    int blo[2], bhi[2], bself[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        blo[r] = win ? lo[r] + shift : 0;
        bhi[r] = win ? hi[r] + shift : 0;
        bself[r] = (win && self[r] >= 0) ? self[r] + shift : -1;
    }
    const int w_end = win ? min(((max(w_hi + shift, kv_len) + MKT - 1) / MKT) * MKT, n_rows) : n1p;

    // ---- main loop (identical in structure to attention_mma.cu)
    const float sl = a.scale * 1.4426950408889634f;
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
            corr[r] = (m_new == -INFINITY) ? 1.f : fast_exp2((m_run[r] - m_new) * sl);
            m_run[r] = m_new;
            msl[r] = (m_new == -INFINITY) ? 0.f : m_new * sl;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = fast_exp2(fmaf(s[nb][0], sl, -msl[0]));
            s[nb][1] = fast_exp2(fmaf(s[nb][1], sl, -msl[0]));
            s[nb][2] = fast_exp2(fmaf(s[nb][2], sl, -msl[1]));
            s[nb][3] = fast_exp2(fmaf(s[nb][3], sl, -msl[1]));
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
        if (row0 < q_len) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row0) * a.ldo + col) = pack2<FP16>(o[i][0] * inv0, o[i][1] * inv0);
        if (row1 < q_len) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row1) * a.ldo + col) = pack2<FP16>(o[i][2] * inv1, o[i][3] * inv1);
    }
}

// ------------------------------------------------------------------------------------------------
// Candidate rows of a unit over  context U own rows  (text self-attention, D = 64): persistent over the
// unit's query tiles.  The context K/V (the big, shared part) is staged ONCE per CTA; per 128-row query tile
// only the Q rows and the window  [tile - halo, tile + 128 + halo)  (halo = longest candidate - 1 rows, so it
// contains every row's own-candidate interval) are loaded, double-buffered with cp.async so that the loads of
// tile i+1 overlap the tensor-core work of tile i.
// ------------------------------------------------------------------------------------------------
template <bool FP16>
__global__ void __launch_bounds__(256)
attn_cand_kernel(AttnJobsArgs a, int halo) {
    constexpr int D = 64, NW = 8, MQT = 128, LD = D + PADE, NT = 256, CH = D / 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* Ks0 = reinterpret_cast<bf16*>(smem_raw);                       // [kv_cap][LD]  context keys
    bf16* Vs0 = Ks0 + static_cast<size_t>(a.kv_cap) * LD;               // [kv_cap][LD]
    bf16* buf0 = Vs0 + static_cast<size_t>(a.kv_cap) * LD;              // 2 x { Q [MQT][LD], Kw [win_cap][LD], Vw [win_cap][LD] }
    const size_t buf_elems = static_cast<size_t>(MQT + 2 * a.win_cap) * LD;

    const int* job = a.jobs + static_cast<size_t>(blockIdx.z) * 8;
    const int q_start = job[0], q_len = job[1], kv_start = job[2], kv_len = job[3];
    const int h = blockIdx.y;
    const int n_tiles = (q_len + MQT - 1) / MQT;
    if (static_cast<int>(blockIdx.x) >= n_tiles) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const bf16* Q = static_cast<const bf16*>(a.q) + h * D;
    const bf16* K = static_cast<const bf16*>(a.k) + h * D;
    const bf16* V = static_cast<const bf16*>(a.v) + h * D;
    bf16* O = static_cast<bf16*>(a.o) + h * D;
    const int n1p = ((kv_len + MKT - 1) / MKT) * MKT;
    const int q_end = q_start + q_len;                                   // packed-row end of the job

    // context K/V: once
    for (int i = tid; i < n1p * CH; i += NT) {
        const int r = i / CH, c = (i % CH) * 8;
        if (r < kv_len) {
            cp_async16(Ks0 + r * LD + c, K + static_cast<size_t>(kv_start + r) * a.ldk + c);
            cp_async16(Vs0 + r * LD + c, V + static_cast<size_t>(kv_start + r) * a.ldv + c);
        } else {
            *reinterpret_cast<uint4*>(Ks0 + r * LD + c) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(Vs0 + r * LD + c) = make_uint4(0, 0, 0, 0);
        }
    }
    auto window_of = [&](int tile, int& w_lo, int& n2) {
        const int r0 = q_start + tile * MQT;
        w_lo = max(q_start, r0 - halo);
        n2 = min(min(q_end, r0 + MQT + halo) - w_lo, a.win_cap);
    };
    auto prefetch = [&](int tile, int b) {
        bf16* Qs = buf0 + b * buf_elems;
        bf16* Kw = Qs + MQT * LD;
        bf16* Vw = Kw + static_cast<size_t>(a.win_cap) * LD;
        const int r0 = q_start + tile * MQT;
        for (int i = tid; i < MQT * CH; i += NT) {
            const int r = i / CH, c = (i % CH) * 8;
            if (r0 + r < q_end) cp_async16(Qs + r * LD + c, Q + static_cast<size_t>(r0 + r) * a.ldq + c);
            else *reinterpret_cast<uint4*>(Qs + r * LD + c) = make_uint4(0, 0, 0, 0);
        }
        int w_lo, n2;
        window_of(tile, w_lo, n2);
        const int n2p = ((n2 + MKT - 1) / MKT) * MKT;
        for (int i = tid; i < n2p * CH; i += NT) {
            const int r = i / CH, c = (i % CH) * 8;
            if (r < n2) {
                cp_async16(Kw + r * LD + c, K + static_cast<size_t>(w_lo + r) * a.ldk + c);
                cp_async16(Vw + r * LD + c, V + static_cast<size_t>(w_lo + r) * a.ldv + c);
            } else {
                *reinterpret_cast<uint4*>(Kw + r * LD + c) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(Vw + r * LD + c) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const float sl = a.scale * 1.4426950408889634f;
    int b = 0;
    prefetch(blockIdx.x, 0);                                             // group 0 = context K/V + first tile
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, b ^= 1) {
        const int next = tile + gridDim.x;
        if (next < n_tiles) {
            prefetch(next, b ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");       // everything but the newest group has landed
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        // this thread's two rows: own-candidate intervals (global loads overlap the barrier below)
        const int r0 = q_start + tile * MQT;
        const int ra = min(r0 + warp * 16 + g, q_end - 1), rb = min(r0 + warp * 16 + g + 8, q_end - 1);
        const int4 ia = *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(ra) * 4);
        const int4 ib = *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(rb) * 4);
        __syncthreads();
        const bf16* Qs = buf0 + b * buf_elems;
        const bf16* Kw = Qs + MQT * LD;
        const bf16* Vw = Kw + static_cast<size_t>(a.win_cap) * LD;
        int w_lo, n2;
        window_of(tile, w_lo, n2);
        // window-buffer coordinates of the two rows' intervals, and the warp's window tile range
        const int lo0 = ia.x - w_lo, hi0 = ia.y - w_lo, sf0 = ia.z >= 0 ? ia.z - w_lo : -1;
        const int lo1 = ib.x - w_lo, hi1 = ib.y - w_lo, sf1 = ib.z >= 0 ? ib.z - w_lo : -1;
        int wmin = min(min(lo0, lo1), min(sf0 >= 0 ? sf0 : lo0, sf1 >= 0 ? sf1 : lo1));
        int wmax = max(max(hi0, hi1), max(sf0 + 1, sf1 + 1));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wmin = min(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
            wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        }
        const int wt_begin = (max(wmin, 0) / MKT) * MKT;
        const int wt_end = min(((max(wmax, 0) + MKT - 1) / MKT) * MKT, ((n2 + MKT - 1) / MKT) * MKT);

        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
        float o[D / 8][4];
#pragma unroll
        for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        const bf16* q_base = Qs + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (lane >> 4);

        auto process = [&](const bf16* k_tile, const bf16* v_tile, unsigned long long m0, unsigned long long m1) {
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
                corr[r] = (m_new == -INFINITY) ? 1.f : fast_exp2((m_run[r] - m_new) * sl);
                m_run[r] = m_new;
                msl[r] = (m_new == -INFINITY) ? 0.f : m_new * sl;
                l_run[r] *= corr[r];
            }
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                s[nb][0] = fast_exp2(fmaf(s[nb][0], sl, -msl[0]));
                s[nb][1] = fast_exp2(fmaf(s[nb][1], sl, -msl[0]));
                s[nb][2] = fast_exp2(fmaf(s[nb][2], sl, -msl[1]));
                s[nb][3] = fast_exp2(fmaf(s[nb][3], sl, -msl[1]));
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
        };
        if (r0 + warp * 16 < q_end) {                                    // warps past the job's last row have nothing to do
            for (int t0 = 0; t0 < n1p; t0 += MKT) {                      // context: every real key allowed
                const unsigned long long m = tile_mask(0, kv_len, -1, t0);
                process(Ks0 + static_cast<size_t>(t0) * LD, Vs0 + static_cast<size_t>(t0) * LD, m, m);
            }
            for (int t0 = wt_begin; t0 < wt_end; t0 += MKT)              // own candidate: interval + self
                process(Kw + static_cast<size_t>(t0) * LD, Vw + static_cast<size_t>(t0) * LD, tile_mask(lo0, hi0, sf0, t0),
                        tile_mask(lo1, hi1, sf1, t0));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
                l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
            }
            const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
            const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
            const int row0 = r0 + warp * 16 + g, row1 = row0 + 8;
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                const int col = i * 8 + 2 * t;
                if (row0 < q_end) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row0) * a.ldo + col) = pack2<FP16>(o[i][0] * inv0, o[i][1] * inv0);
                if (row1 < q_end) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row1) * a.ldo + col) = pack2<FP16>(o[i][2] * inv1, o[i][3] * inv1);
            }
        }
        __syncthreads();   // buffer b is refilled by the prefetch issued in the next iteration
    }
}

template <bool FP16>
int launch_cand(const AttnJobsArgs& a, int halo, cudaStream_t stream) {
    const size_t smem = sizeof(bf16) * (2 * static_cast<size_t>(a.kv_cap) + 2 * (128 + 2 * static_cast<size_t>(a.win_cap))) * (64 + PADE);
    UNIMM_CHECK(smem <= 227 * 1024, "candidate attention: staging does not fit shared memory");
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_cand_kernel<FP16>), smem));
    const int max_tiles = (a.max_q_len + 127) / 128;
    dim3 grid(max_tiles < 2 ? max_tiles : 2, a.heads, a.n_jobs);     // two CTAs share a unit's tiles: ~2 x heads x units CTAs
    attn_cand_kernel<FP16><<<grid, 256, smem, stream>>>(a, halo);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core version of the same job semantics (fp32 parity mode; simple, one warp per query row)
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
attn_jobs_simt_kernel(AttnJobsArgs a) {
    const int* job = a.jobs + static_cast<size_t>(blockIdx.z) * 8;
    const int q_start = job[0], q_len = job[1], kv_start = job[2], kv_len = job[3], win = job[4], mask_row = job[5];
    const int h = blockIdx.y, lane = threadIdx.x & 31;
    const int qr = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qr >= q_len) return;
    const float* Q = static_cast<const float*>(a.q) + static_cast<size_t>(q_start + qr) * a.ldq + h * D;
    const float* K = static_cast<const float*>(a.k) + h * D;
    const float* V = static_cast<const float*>(a.v) + h * D;
    float* O = static_cast<float*>(a.o) + static_cast<size_t>(q_start + qr) * a.ldo + h * D;
    int lo = 0, hi = 0, self = -1;
    if (win) {
        const int4 iv = *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(q_start + qr) * 4);
        lo = iv.x; hi = iv.y; self = iv.z;
    }
    const float* km = mask_row >= 0 ? a.key_mask + static_cast<size_t>(mask_row) * a.key_mask_ld : nullptr;
    bool any_key = true;
    if (km != nullptr) {
        int any = 0;
        for (int k = lane; k < kv_len; k += 32) any |= (km[k] > 0.5f);
        any_key = __any_sync(0xffffffffu, any);
    }
    float q[D / 32];
#pragma unroll
    for (int i = 0; i < D / 32; ++i) q[i] = Q[lane + 32 * i];
    float m = -INFINITY, l = 0.f, o[D / 32];
#pragma unroll
    for (int i = 0; i < D / 32; ++i) o[i] = 0.f;
    // key list: shared range, then the row's own interval, then itself
    const int n_keys = kv_len + (hi > lo ? hi - lo : 0) + (self >= 0 ? 1 : 0);
    for (int j = 0; j < n_keys; ++j) {
        int src;
        if (j < kv_len) {
            if (km != nullptr && any_key && !(km[j] > 0.5f)) continue;
            src = kv_start + j;
        } else if (j < kv_len + (hi > lo ? hi - lo : 0)) {
            src = lo + (j - kv_len);
        } else {
            src = self;
        }
        const float* kr = K + static_cast<size_t>(src) * a.ldk;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) s = fmaf(q[i], kr[lane + 32 * i], s);
        s = warp_sum(s) * a.scale;
        const float m_new = fmaxf(m, s);
        const float corr = expf(m - m_new), p = expf(s - m_new);
        l = l * corr + p;
        const float* vr = V + static_cast<size_t>(src) * a.ldv;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) o[i] = o[i] * corr + p * vr[lane + 32 * i];
        m = m_new;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) O[lane + 32 * i] = o[i] * inv;
}

template <int D, bool FP16, int NW>
int launch_jobs(const AttnJobsArgs& a, cudaStream_t stream) {
    constexpr int MQT = 16 * NW;
    const size_t smem = sizeof(bf16) * (MQT + 2 * static_cast<size_t>(a.kv_cap + a.win_cap)) * (D + PADE);
    UNIMM_CHECK(smem <= 227 * 1024, "attention jobs: staged key range does not fit shared memory");
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_jobs_kernel<D, FP16, NW>), smem));
    dim3 grid((a.max_q_len + MQT - 1) / MQT, a.heads, a.n_jobs);
    attn_jobs_kernel<D, FP16, NW><<<grid, NW * 32, smem, stream>>>(a);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

template <int D, bool FP16>
int dispatch_jobs(const AttnJobsArgs& a, cudaStream_t stream) {
    return a.max_q_len > 64 ? launch_jobs<D, FP16, 8>(a, stream) : launch_jobs<D, FP16, 4>(a, stream);
}

}  // namespace

// candidate jobs (win = 1 for every job, D = 64) through the persistent double-buffered kernel
int attention_candidates(const AttnJobsArgs& a, int halo, cudaStream_t stream) {
    UNIMM_CHECK(a.n_jobs > 0 && a.n_jobs <= 65535 && a.D == 64 && a.win_cap % 64 == 0 && a.kv_cap % 64 == 0 && a.kv_cap <= 256,
                "candidate attention: bad arguments");
    UNIMM_CHECK(a.win_cap >= 128 + 2 * halo, "candidate attention: window capacity smaller than tile + 2 * halo");
    UNIMM_CHECK((a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 2) == 0, "attention jobs: rows must be 16-byte aligned");
    return a.lp_kind == LP_FP16 ? launch_cand<true>(a, halo, stream) : launch_cand<false>(a, halo, stream);
}

int attention_jobs(const AttnJobsArgs& a, bool fp32, cudaStream_t stream) {
    UNIMM_CHECK(a.n_jobs > 0 && a.n_jobs <= 65535 && a.heads > 0 && a.max_q_len > 0, "attention jobs: bad problem size");
    UNIMM_CHECK(a.D == 64 || a.D == 128, "attention jobs: head dim must be 64 or 128");
    UNIMM_CHECK(a.kv_cap % 64 == 0 && a.win_cap % 64 == 0 && a.kv_cap > 0 && a.kv_cap <= 256, "attention jobs: bad staging capacity");
    if (fp32) {
        dim3 grid((a.max_q_len + 3) / 4, a.heads, a.n_jobs);
        if (a.D == 64) attn_jobs_simt_kernel<64><<<grid, 128, 0, stream>>>(a);
        else attn_jobs_simt_kernel<128><<<grid, 128, 0, stream>>>(a);
        UNIMM_LAUNCH_CHECK(1);
        return 0;
    }
    UNIMM_CHECK((a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 2) == 0, "attention jobs: rows must be 16-byte aligned");
    if (a.lp_kind == LP_FP16) return a.D == 64 ? dispatch_jobs<64, true>(a, stream) : dispatch_jobs<128, true>(a, stream);
    return a.D == 64 ? dispatch_jobs<64, false>(a, stream) : dispatch_jobs<128, false>(a, stream);
}

}  // namespace unimm
