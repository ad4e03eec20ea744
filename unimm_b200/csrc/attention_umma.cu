// tcgen05 / TMEM / TMA attention for the CANDIDATE rows of the prefix-shared layout (text self-attention, D = 64).
//
// Replaces, for the rows a candidate owns, the reference's matmul / +mask / softmax / matmul of BertSelfAttention
// (models/vilbert_dialog.py:395-410) under the generative mask of utils/data_utils.py:199-210.  A candidate row r
// attends   context rows [kv_start, kv_start + kv_len)  U  own-candidate rows [lo_r, hi_r) U {self_r}   (attention_jobs.cu).
// 94 % of those (query, key) pairs are the context part, which is the same dense [q_len x kv_len] problem for every
// row of a unit, so it runs on the 5th-generation tensor cores; the <= 17 own keys of a row lie within +-16 packed
// rows of it and are handled by the softmax warps themselves with mma.sync on a TMA-staged window.  The two partial
// softmaxes (each with its own running max) are merged in the epilogue.
//
// One persistent CTA per SM walks (unit, head) ITEMS; per item the context K / V (<= 256 keys) sit in shared memory and
// the item's query rows stream through in 128-row tiles:
//
//   warp 0   TMA producer: Q tile (2 boxes of 64 rows x 128 B, SWIZZLE_128B) and the K / V window rows
//            [tile - 16, tile + 176) into 2-deep rings
//   warp 3   TMA producer of the per-item context K and V (single buffers, refilled as soon as the last QK / PV MMA of
//            the previous item has read them: the K refill overlaps the previous item's last softmax)
//   warp 1   one thread issues  S[b] = Q K_ctx^T   (tcgen05.mma M128 x N(64..256) x K16, both operands K-major in smem,
//            fp32 accumulator = TMEM columns [256 b, 256 b + N))      one tile AHEAD of
//                            O[b] = P[b] V_ctx    (A = P from TMEM, B = V as stored = MN-major in smem, N = 64)
//   warp 2   TMEM allocator (512 columns: two S / P / O buffers)
//   warps 4-7 / 8-11   two softmax warpgroups on alternating tiles (thread = one query row = one TMEM lane):
//            pass 1 row max, pass 2 p = 2^(s*c - m*c) -> 16-bit pairs written back over S with tcgen05.st (P never
//            touches shared memory), own-candidate part with mma.sync while the PV MMA runs, then O from TMEM in the
//            mma fragment arrangement (tcgen05.ld.16x256b), merge, normalise, 16-bit stores.
#include <cuda.h>

#include <cstdlib>

#include "attn_common.cuh"
#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace unimm {
namespace {

using namespace attn;

constexpr int TQ = 128;                 // query rows per tile (= UMMA M)
constexpr int BOXR = 64;                // rows per TMA box
constexpr int BOXB = BOXR * 128;        // bytes per box: 64 rows x 64 16-bit elements
constexpr int HALO = 16;                // own-candidate keys of row r lie in [r - HALO, r + HALO]
constexpr int WBOXR = 32;               // rows per window TMA box
constexpr int WBOXB = WBOXR * 128;
constexpr int WBOX = 5;                 // window boxes: rows [tile - 16, tile + 144) = tile + 2 * HALO
constexpr int WSTAGE = WBOX * WBOXB;    // 20 KB
constexpr int KC_OFF = 0;               // context K: 4 boxes
constexpr int VC_OFF = 4 * BOXB;        // context V: 4 boxes
constexpr int Q_OFF = 8 * BOXB;         // 2 x 2 boxes
constexpr int KW_OFF = 12 * BOXB;       // 2 stages
constexpr int VW_OFF = KW_OFF + 2 * WSTAGE;
constexpr int OST_OFF = VW_OFF + 2 * WSTAGE;   // output staging: 8 warps x 32 rows x 128 B
constexpr int BAR_OFF = OST_OFF + 8 * 4096;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024 /* alignment slack */;
constexpr int O_COL = 192;              // O accumulator columns inside a 256-column TMEM buffer (P uses [0,128))

enum Bar { KC_FULL = 0, KC_FREE, VC_FULL, VC_FREE, Q_FULL, Q_FREE = Q_FULL + 2, W_FULL = Q_FREE + 2, W_FREE = W_FULL + 2,
           S_FULL = W_FREE + 2, S_FREE = S_FULL + 2, P_FULL = S_FREE + 2, O_FULL = P_FULL + 2, NBAR = O_FULL + 2 };

// byte offset of 16-byte chunk `chunk` of row `row` inside a stack of SWIZZLE_128B boxes (rows of 128 bytes)
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// polling wait of the single-thread roles (TMA producers, MMA issuer): back off between probes so that the spin does not
// take issue slots from the softmax warps that share the scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!ptx::mbar_try_wait(bar, parity)) {
        __nanosleep(40);
        if (++spins > (1u << 24)) __trap();
    }
}

struct Item {
    int q_start, q_len, kv_start, kv_len, head, n_tiles, nkb, job;
};
__device__ __forceinline__ Item load_item(const AttnJobsArgs& a, int item) {
    Item it;
    const int job = item / a.heads;
    it.head = item - job * a.heads;
    it.job = job;
    if (a.jobs == nullptr) {      // dense layout: job = sequence, its seq_len rows are queries and keys
        it.q_start = job * a.seq_len; it.q_len = a.seq_len; it.kv_start = it.q_start; it.kv_len = a.seq_len;
    } else {
        const int4 j = *reinterpret_cast<const int4*>(a.jobs + static_cast<size_t>(job) * 8);
        it.q_start = j.x; it.q_len = j.y; it.kv_start = j.z; it.kv_len = j.w;
    }
    it.n_tiles = (it.q_len + TQ - 1) / TQ;
    it.nkb = (it.kv_len + BOXR - 1) / BOXR;
    return it;
}

// DENSE = true: the dense [B, S <= 256] layout (discriminative scoring, training forward): item = (sequence, head), the
// sequence's own S rows are the "context" keys, the allowed columns of a row come from its descriptor (text_row_interval:
// generative / discriminative masks of utils/data_utils.py:149-210, :300-354; padding rows attend everything, exactly like the
// reference's additive -10000), and there is no own-candidate part and no window ring.
template <bool FP16, bool DENSE>
__global__ void __launch_bounds__(384, 1)
attn_cand_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKw,
                      const __grid_constant__ CUtensorMap tmVw, AttnJobsArgs a, int n_items, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmK);
        ptx::prefetch_tensormap(&tmV);
        ptx::prefetch_tensormap(&tmKw);
        ptx::prefetch_tensormap(&tmVw);
    }
    if (warp == 1 && ptx::elect_one()) {
        ptx::mbar_init(&bars[KC_FULL], 1); ptx::mbar_init(&bars[KC_FREE], 1);
        ptx::mbar_init(&bars[VC_FULL], 1); ptx::mbar_init(&bars[VC_FREE], 1);
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&bars[Q_FULL + b], 1); ptx::mbar_init(&bars[Q_FREE + b], 4);
            ptx::mbar_init(&bars[W_FULL + b], 1); ptx::mbar_init(&bars[W_FREE + b], 4);
            ptx::mbar_init(&bars[S_FULL + b], 1); ptx::mbar_init(&bars[S_FREE + b], 4);
            ptx::mbar_init(&bars[P_FULL + b], 4); ptx::mbar_init(&bars[O_FULL + b], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = ptx::smem_u32(smem);
    ptx::grid_dep_launch_dependents();
    ptx::grid_dep_wait();                  // the prologue above overlapped the previous kernel's tail; Q / K / V are its outputs

    if (warp == 0) {
        // ---------------------------------------------------------------- Q / window producer
        if (ptx::elect_one()) {
            int n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const Item it = load_item(a, item);
                const int col = it.head * 64;
                for (int t = 0; t < it.n_tiles; ++t, ++n) {
                    const int b = n & 1;
                    const uint32_t ph = (n >> 1) & 1;
                    const int r0 = it.q_start + t * TQ;
                    ptx::mbar_wait(&bars[Q_FREE + b], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&bars[Q_FULL + b], 2 * BOXB);
                    ptx::tma_load_2d(smem + Q_OFF + b * 2 * BOXB, &tmQ, &bars[Q_FULL + b], col, r0);
                    ptx::tma_load_2d(smem + Q_OFF + b * 2 * BOXB + BOXB, &tmQ, &bars[Q_FULL + b], col, r0 + BOXR);
                    if (!DENSE) {
                        ptx::mbar_wait(&bars[W_FREE + b], ph ^ 1);
                        ptx::mbar_arrive_expect_tx(&bars[W_FULL + b], 2 * WSTAGE);
#pragma unroll
                        for (int j = 0; j < WBOX; ++j) {
                            ptx::tma_load_2d(smem + KW_OFF + b * WSTAGE + j * WBOXB, &tmKw, &bars[W_FULL + b], col, r0 - HALO + j * WBOXR);
                            ptx::tma_load_2d(smem + VW_OFF + b * WSTAGE + j * WBOXB, &tmVw, &bars[W_FULL + b], col, r0 - HALO + j * WBOXR);
                        }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ---------------------------------------------------------------- context K / V producer
        if (ptx::elect_one()) {
            uint32_t k = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
                const Item it = load_item(a, item);
                const int col = it.head * 64;
                ptx::mbar_wait(&bars[KC_FREE], (k & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&bars[KC_FULL], it.nkb * BOXB);
                for (int j = 0; j < it.nkb; ++j)
                    ptx::tma_load_2d(smem + KC_OFF + j * BOXB, &tmK, &bars[KC_FULL], col, it.kv_start + j * BOXR);
                ptx::mbar_wait(&bars[VC_FREE], (k & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&bars[VC_FULL], it.nkb * BOXB);
                for (int j = 0; j < it.nkb; ++j)
                    ptx::tma_load_2d(smem + VC_OFF + j * BOXB, &tmV, &bars[VC_FULL], col, it.kv_start + j * BOXR);
            }
            // the last item's "free" commits must have landed in this CTA's shared memory before it exits
            if (k > 0) {
                ptx::mbar_wait(&bars[KC_FREE], (k - 1) & 1);
                ptx::mbar_wait(&bars[VC_FREE], (k - 1) & 1);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer: QK(n + 1) ahead of PV(n)
        if (ptx::elect_one()) {
            const uint32_t fmt = FP16 ? 0u : 1u;
            struct Cursor { int item; uint32_t seq; int t; int n; Item it; bool valid; };
            auto start = [&](Cursor& c) {
                c.item = blockIdx.x; c.seq = 0; c.t = 0; c.n = 0;
                c.valid = c.item < n_items;
                if (c.valid) c.it = load_item(a, c.item);
            };
            auto advance = [&](Cursor& c) {
                ++c.n;
                if (++c.t == c.it.n_tiles) {
                    c.t = 0; ++c.seq; c.item += gridDim.x;
                    c.valid = c.item < n_items;
                    if (c.valid) c.it = load_item(a, c.item);
                }
            };
            auto issue_qk = [&](const Cursor& c) {
                const int b = c.n & 1;
                const uint32_t ph = (c.n >> 1) & 1;
                if (c.t == 0) ptx::mbar_wait(&bars[KC_FULL], c.seq & 1);
                ptx::mbar_wait(&bars[Q_FULL + b], ph);
                ptx::mbar_wait(&bars[S_FREE + b], ph ^ 1);
                ptx::tc_fence_after();
                const uint32_t idesc = ptx::make_idesc_f16(TQ, c.it.nkb * BOXR, fmt);
                const uint64_t da = ptx::make_sw128_kmajor_desc(smem_base + Q_OFF + b * 2 * BOXB);
                const uint64_t db = ptx::make_sw128_kmajor_desc(smem_base + KC_OFF);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) ptx::umma_f16_ss(tmem_base + b * 256, da + 2 * ks, db + 2 * ks, idesc, ks != 0 ? 1u : 0u);
                ptx::umma_commit(&bars[S_FULL + b]);
                if (c.t == c.it.n_tiles - 1) ptx::umma_commit(&bars[KC_FREE]);
            };
            auto issue_pv = [&](const Cursor& c) {
                const int b = c.n & 1;
                const uint32_t ph = (c.n >> 1) & 1;
                if (c.t == 0) ptx::mbar_wait(&bars[VC_FULL], c.seq & 1);
                ptx::mbar_wait(&bars[P_FULL + b], ph);
                ptx::tc_fence_after();
                const uint32_t idesc = ptx::make_idesc_f16(TQ, 64, fmt) | (1u << 16);   // B (= V as stored) is MN-major
                const uint64_t db = ptx::make_sw128_kmajor_desc(smem_base + VC_OFF);    // 8-key groups 1024 B apart
                const int nks = c.it.nkb * 4;                                           // 16 keys per MMA
                for (int ks = 0; ks < nks; ++ks)
                    ptx::umma_f16_ts(tmem_base + b * 256 + O_COL, tmem_base + b * 256 + 8 * ks, db + 128 * ks, idesc, ks != 0 ? 1u : 0u);
                ptx::umma_commit(&bars[O_FULL + b]);
                if (c.t == c.it.n_tiles - 1) ptx::umma_commit(&bars[VC_FREE]);
            };
            Cursor cq, cp;
            start(cq);
            start(cp);
            if (cq.valid) { issue_qk(cq); advance(cq); }
            while (cp.valid) {
                if (cq.valid) { issue_qk(cq); advance(cq); }
                issue_pv(cp);
                advance(cp);
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- softmax / own-candidate / epilogue warpgroups
        const int wg = (warp - 4) >> 2;
        const int qd = warp & 3;                                   // TMEM lane quarter this warp may touch
        const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
        const int g = lane >> 2, t4 = lane & 3;
        const float sl = a.scale * 1.4426950408889634f;
        int n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Item it = load_item(a, item);
            const int n1p = it.nkb * BOXR;
            const int q_end = it.q_start + it.q_len;
            bf16* Og = static_cast<bf16*>(a.o) + it.head * 64;
            for (int t = 0; t < it.n_tiles; ++t, ++n) {
                if ((n & 1) != wg) continue;
                const int b = wg;
                const uint32_t ph = (n >> 1) & 1;
                const int r0 = it.q_start + t * TQ;
                const uint32_t tS = tmem_base + b * 256 + lane_off;
                const uint32_t ost = smem_base + OST_OFF + (warp - 4) * 4096;   // this warp's 32 rows x 128 B stash / staging block

                // allowed context columns of this thread's row: [c_lo, c_hi) U {c_self}
                int c_lo = 0, c_hi = it.kv_len, c_self = -1;
                if (DENSE) {
                    const int r = t * TQ + qd * 32 + lane;
                    text_row_interval(a.desc[it.job], r, a.seq_len, c_lo, c_hi, c_self);
                    if (c_hi <= c_lo && c_self < 0) { c_lo = 0; c_hi = a.seq_len; }      // padding row: softmax over the raw scores
                }
                ptx::mbar_wait(&bars[S_FULL + b], ph);
                ptx::tc_fence_after();

                // ---- (1) pass 1: row max over the context keys (the next 32 columns are in flight while 32 are folded)
                float mx = -INFINITY;
                uint32_t v0[32], v1[32];
                auto fold_max = [&](const uint32_t* v, int c) {
                    if (c >= c_lo && c + 32 <= c_hi) {
                        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
                            m2 = fmaxf(m2, __uint_as_float(v[j + 2])); m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
                        }
                        mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if ((c + j >= c_lo && c + j < c_hi) || c + j == c_self) mx = fmaxf(mx, __uint_as_float(v[j]));
                    }
                };
                if (dbg & 2) mx = 0.f;
                else {
                    ptx::tmem_ld_32x32b_x32(tS, v0);
                    for (int c = 0; c < n1p; c += 64) {
                        ptx::tmem_ld_wait();
                        ptx::tmem_ld_32x32b_x32(tS + c + 32, v1);
                        fold_max(v0, c);
                        ptx::tmem_ld_wait();
                        if (c + 64 < n1p) ptx::tmem_ld_32x32b_x32(tS + c + 64, v0);
                        fold_max(v1, c + 32);
                    }
                }
                // own-candidate intervals of this thread's 4 fragment rows (global loads fly during pass 2)
                int4 iv[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int r = min(r0 + qd * 32 + mt * 16 + g + 8 * hh, q_end - 1);
                        iv[mt][hh] = DENSE ? make_int4(0, 0, -1, 0) : *reinterpret_cast<const int4*>(a.row_iv + static_cast<size_t>(r) * 4);
                    }
                const float m_ctx = mx;
                const float msl = m_ctx * sl;

                // ---- (2) pass 2: p = 2^(s*sl - m*sl), 16-bit pairs written back over S (columns [0, n1p / 2))
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                const bool poly = (dbg & 16) != 0;      // UNIMM_ATTN_DBG=16: half of the exponentials on the FMA pipe (measured SLOWER: DESIGN.md 5c)
                auto emit_p = [&](const uint32_t* v, int c) {
                    uint32_t pk[16];
                    const bool full = c >= c_lo && c + 32 <= c_hi;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float p0 = fast_exp2(fmaf(__uint_as_float(v[j]), sl, -msl));
                        float p1 = fast_exp2(fmaf(__uint_as_float(v[j + 1]), sl, -msl));
                        float p2, p3;
                        if (poly) {                     // half of the exponentials off the SFU (exp2_poly2, attn_common.cuh)
                            exp2_poly2(fmaf(__uint_as_float(v[j + 2]), sl, -msl), fmaf(__uint_as_float(v[j + 3]), sl, -msl), p2, p3);
                        } else {
                            p2 = fast_exp2(fmaf(__uint_as_float(v[j + 2]), sl, -msl));
                            p3 = fast_exp2(fmaf(__uint_as_float(v[j + 3]), sl, -msl));
                        }
                        if (!full) {
                            const int cj = c + j;
                            if (!((cj >= c_lo && cj < c_hi) || cj == c_self)) p0 = 0.f;
                            if (!((cj + 1 >= c_lo && cj + 1 < c_hi) || cj + 1 == c_self)) p1 = 0.f;
                            if (!((cj + 2 >= c_lo && cj + 2 < c_hi) || cj + 2 == c_self)) p2 = 0.f;
                            if (!((cj + 3 >= c_lo && cj + 3 < c_hi) || cj + 3 == c_self)) p3 = 0.f;
                        }
                        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
                        pk[j >> 1] = pack2<FP16>(p0, p1);
                        pk[(j >> 1) + 1] = pack2<FP16>(p2, p3);
                    }
                    tmem_st_32x32b_x16(tS + (c >> 1), pk);
                };
                if (!(dbg & 4)) {
                    ptx::tmem_ld_32x32b_x32(tS, v0);
                    for (int c = 0; c < n1p; c += 64) {
                        ptx::tmem_ld_wait();
                        ptx::tmem_ld_32x32b_x32(tS + c + 32, v1);
                        emit_p(v0, c);
                        ptx::tmem_ld_wait();
                        if (c + 64 < n1p) ptx::tmem_ld_32x32b_x32(tS + c + 64, v0);
                        emit_p(v1, c + 32);
                    }
                }
                const float l_ctx = (l0 + l1) + (l2 + l3);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars[P_FULL + b]);

                // ---- (3) own-candidate part with mma.sync while the PV MMA runs: keys = window rows [rowbase, rowbase + 48)
                // rows without own keys (context rows scored in the same launch, the dense layout): nothing to do for the warp
                bool any_own = false;
                if (!DENSE) {
                    bool mine = false;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) mine |= iv[mt][hh].y > iv[mt][hh].x || iv[mt][hh].z >= 0;
                    any_own = __any_sync(0xffffffffu, mine) && !(dbg & 1);
                    ptx::mbar_wait(&bars[Q_FULL + b], ph);
                    ptx::mbar_wait(&bars[W_FULL + b], ph);
                }
                const uint32_t qs = smem_base + Q_OFF + b * 2 * BOXB;
                const uint32_t kw = smem_base + KW_OFF + b * WSTAGE;
                const uint32_t vw = smem_base + VW_OFF + b * WSTAGE;
                float o_own[2][8][4];
                float m_own[2][2], l_own[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (!any_own) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) o_own[mt][i][0] = o_own[mt][i][1] = o_own[mt][i][2] = o_own[mt][i][3] = 0.f;
                        m_own[mt][0] = m_own[mt][1] = -INFINITY; l_own[mt][0] = l_own[mt][1] = 0.f;
                        continue;
                    }
                    const int rowbase = qd * 32 + mt * 16;              // keys = window rows [rowbase, rowbase + 48)
                    float s[6][4];
#pragma unroll
                    for (int i = 0; i < 6; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t qa[4];
                        ldsm4(qa, qs + sw128(rowbase + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * ks + (lane >> 4)));
#pragma unroll
                        for (int nb2 = 0; nb2 < 3; ++nb2) {
                            uint32_t kb[4];
                            ldsm4(kb, kw + sw128(rowbase + nb2 * 16 + (lane & 7) + 8 * (lane >> 4), 2 * ks + ((lane >> 3) & 1)));
                            mma_lp<FP16>(s[2 * nb2], qa, kb[0], kb[1]);
                            mma_lp<FP16>(s[2 * nb2 + 1], qa, kb[2], kb[3]);
                        }
                    }
                    const int key0 = r0 - HALO + rowbase + 2 * t4;          // packed row of this thread's first key column
                    float tm[2] = {-INFINITY, -INFINITY};
#pragma unroll
                    for (int nb = 0; nb < 6; ++nb)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int hh = e >> 1, key = key0 + nb * 8 + (e & 1);
                            const bool ok = (key >= iv[mt][hh].x && key < iv[mt][hh].y) || key == iv[mt][hh].z;
                            s[nb][e] = ok ? s[nb][e] : -INFINITY;
                            tm[hh] = fmaxf(tm[hh], s[nb][e]);
                        }
                    float ls[2] = {0.f, 0.f};
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        tm[hh] = fmaxf(tm[hh], __shfl_xor_sync(0xffffffffu, tm[hh], 1));
                        tm[hh] = fmaxf(tm[hh], __shfl_xor_sync(0xffffffffu, tm[hh], 2));
                        m_own[mt][hh] = tm[hh];
                    }
                    const float ms0 = tm[0] == -INFINITY ? 0.f : tm[0] * sl, ms1 = tm[1] == -INFINITY ? 0.f : tm[1] * sl;
#pragma unroll
                    for (int nb = 0; nb < 6; ++nb) {
                        s[nb][0] = fast_exp2(fmaf(s[nb][0], sl, -ms0));
                        s[nb][1] = fast_exp2(fmaf(s[nb][1], sl, -ms0));
                        s[nb][2] = fast_exp2(fmaf(s[nb][2], sl, -ms1));
                        s[nb][3] = fast_exp2(fmaf(s[nb][3], sl, -ms1));
                        ls[0] += s[nb][0] + s[nb][1];
                        ls[1] += s[nb][2] + s[nb][3];
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        ls[hh] += __shfl_xor_sync(0xffffffffu, ls[hh], 1);
                        ls[hh] += __shfl_xor_sync(0xffffffffu, ls[hh], 2);
                        l_own[mt][hh] = ls[hh];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) o_own[mt][i][0] = o_own[mt][i][1] = o_own[mt][i][2] = o_own[mt][i][3] = 0.f;
#pragma unroll
                    for (int kc = 0; kc < 3; ++kc) {
                        uint32_t pa[4];
                        pa[0] = pack2<FP16>(s[2 * kc][0], s[2 * kc][1]);
                        pa[1] = pack2<FP16>(s[2 * kc][2], s[2 * kc][3]);
                        pa[2] = pack2<FP16>(s[2 * kc + 1][0], s[2 * kc + 1][1]);
                        pa[3] = pack2<FP16>(s[2 * kc + 1][2], s[2 * kc + 1][3]);
#pragma unroll
                        for (int db2 = 0; db2 < 4; ++db2) {
                            uint32_t vb[4];
                            ldsm4_trans(vb, vw + sw128(rowbase + kc * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * db2 + (lane >> 4)));
                            mma_lp<FP16>(o_own[mt][2 * db2], pa, vb[0], vb[1]);
                            mma_lp<FP16>(o_own[mt][2 * db2 + 1], pa, vb[2], vb[3]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&bars[Q_FREE + b]);            // (S_FULL was awaited above: the QK MMA has read Q)
                    if (!DENSE) ptx::mbar_arrive(&bars[W_FREE + b]);
                }

                // ---- (4) O = P V from TMEM in the mma fragment arrangement, merge with the stashed own part, normalise
                ptx::mbar_wait(&bars[O_FULL + b], ph);
                ptx::tc_fence_after();
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    float ca[2], cb[2];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int src = mt * 16 + g + 8 * hh;
                        const float mc = __shfl_sync(0xffffffffu, m_ctx, src), lc = __shfl_sync(0xffffffffu, l_ctx, src);
                        const float mo = m_own[mt][hh];
                        const float m = fmaxf(mc, mo);
                        const float ea = fast_exp2((mc - m) * sl);
                        const float eb = mo == -INFINITY ? 0.f : fast_exp2((mo - m) * sl);
                        const float inv = 1.0f / (lc * ea + l_own[mt][hh] * eb);
                        ca[hh] = ea * inv;
                        cb[hh] = eb * inv;
                    }
                    uint32_t u[32];
                    const uint32_t tO = tmem_base + b * 256 + O_COL + lane_off + (static_cast<uint32_t>(mt * 16) << 16);
                    ptx::tmem_ld_16x256b_x4(tO, u);
                    ptx::tmem_ld_16x256b_x4(tO + 32, u + 16);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int lr = mt * 16 + g + 8 * hh;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float x0 = __uint_as_float(u[4 * j + 2 * hh]) * ca[hh] + o_own[mt][j][2 * hh] * cb[hh];
                            const float x1 = __uint_as_float(u[4 * j + 2 * hh + 1]) * ca[hh] + o_own[mt][j][2 * hh + 1] * cb[hh];
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ost + sw128(lr, j) + 4 * t4), "r"(pack2<FP16>(x0, x1)) : "memory");
                        }
                    }
                }
                // O has left TMEM: hand the S / P / O buffer back before the global stores
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars[S_FREE + b]);
                // whole 128-byte rows to global: 8 lanes per row, 4 rows per instruction
                if (!(dbg & 8)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int lr = i * 4 + (lane >> 3), ch = lane & 7;
                        const int row = r0 + qd * 32 + lr;
                        uint32_t x, y, z, w;
                        ptx::lds_v4(ost + sw128(lr, ch), x, y, z, w);
                        if (row < q_end) *reinterpret_cast<uint4*>(Og + static_cast<size_t>(row) * a.ldo + 8 * ch) = make_uint4(x, y, z, w);
                    }
                }
                __syncwarp();
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

template <bool FP16, bool DENSE>
int launch_umma(const AttnJobsArgs& a, cudaStream_t stream) {
    CUtensorMap tmQ, tmK, tmV, tmKw, tmVw;
    const int W = a.heads * 64;
    UNIMM_TRY(gemm_make_map(a.q, a.n_rows, W, a.ldq, BOXR, &tmQ));
    UNIMM_TRY(gemm_make_map(a.k, a.n_rows, W, a.ldk, BOXR, &tmK));
    UNIMM_TRY(gemm_make_map(a.v, a.n_rows, W, a.ldv, BOXR, &tmV));
    UNIMM_TRY(gemm_make_map(a.k, a.n_rows, W, a.ldk, WBOXR, &tmKw));
    UNIMM_TRY(gemm_make_map(a.v, a.n_rows, W, a.ldv, WBOXR, &tmVw));
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_cand_umma_kernel<FP16, DENSE>), SMEM_BYTES));
    const int n_items = a.n_jobs * a.heads;
    const int grid = n_items < gemm_num_sms() ? n_items : gemm_num_sms();
    static int dbg = getenv("UNIMM_ATTN_DBG") ? atoi(getenv("UNIMM_ATTN_DBG")) : 0;   // timing experiments only (results invalid)
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    add_pdl_attr(attr, &cfg.numAttrs);
    UNIMM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_cand_umma_kernel<FP16, DENSE>, tmQ, tmK, tmV, tmKw, tmVw, a, n_items, dbg));
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}


// =================================================================================================================
// Text -> image co-attention of the packed layout (BertBiAttention, models/vilbert_dialog.py:681-698): every text row
// of a unit (context rows and candidate rows are separate jobs) attends the unit's <= 64 image regions, D = 128, keys
// masked by the unit's image mask.  Tiny arithmetic (37 keys) over a large row stream: the kernel is a Q-in / O-out
// stream at HBM speed — Q tiles by TMA, S = Q K^T and O = P V on tcgen05 with S / P / O in TMEM, thread-per-row
// softmax, output rows staged in shared memory and written as whole 256-byte rows.  Same roles as the kernel above.
// =================================================================================================================
namespace x {
constexpr int XD = 128;
constexpr int XKATOM = 64 * 128;          // one 64-column atom of the 64 staged key rows: 8 KB
constexpr int XQATOM = TQ * 128;          // one 64-column atom of a 128-row Q tile: 16 KB
constexpr int XKI_OFF = 0;                // 2 buffers x 2 atoms
constexpr int XVI_OFF = 4 * XKATOM;
constexpr int XQ_OFF = 8 * XKATOM;         // 2 buffers x 2 atoms
constexpr int XOST_OFF = XQ_OFF + 4 * XQATOM;      // 8 warps x 32 rows x 256 B
constexpr int XBAR_OFF = XOST_OFF + 8 * 8192;
constexpr int XSMEM_BYTES = XBAR_OFF + 256 + 1024;
constexpr int XO_COL = 64;                // S / P at [0,64), O at [64,192) of a 256-column buffer
enum Bar { XKV_FULL = 0, XKV_FREE = 2, XQ_FULL = 4, XQ_FREE = 6, XS_FULL = 8, XS_FREE = 10, XP_FULL = 12, XO_FULL = 14, XNBAR = 16 };
}  // namespace x

// SWIZZLE_128B descriptor with an explicit leading-dimension byte offset (distance between 64-element MN atoms of an
// MN-major operand); 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_desc_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

template <bool FP16>
__global__ void __launch_bounds__(384, 1)
attn_cross_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, AttnJobsArgs a, int n_items) {
    using namespace x;   // X-prefixed constants
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XBAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + XNBAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmK);
        ptx::prefetch_tensormap(&tmV);
    }
    if (warp == 1 && ptx::elect_one()) {
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&bars[XKV_FULL + b], 1); ptx::mbar_init(&bars[XKV_FREE + b], 1);
            ptx::mbar_init(&bars[XQ_FULL + b], 1); ptx::mbar_init(&bars[XQ_FREE + b], 1);
            ptx::mbar_init(&bars[XS_FULL + b], 1); ptx::mbar_init(&bars[XS_FREE + b], 4);
            ptx::mbar_init(&bars[XP_FULL + b], 4); ptx::mbar_init(&bars[XO_FULL + b], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem_base = ptx::smem_u32(smem);
    ptx::grid_dep_launch_dependents();
    ptx::grid_dep_wait();                  // the prologue above overlapped the previous kernel's tail; Q / K / V are its outputs

    auto item_of = [&](int item, int& q_start, int& q_len, int& kv_start, int& kv_len, int& mask_row, int& head) {
        const int job = item / a.heads;
        head = item - job * a.heads;
        const int* j = a.jobs + static_cast<size_t>(job) * 8;
        q_start = j[0]; q_len = j[1]; kv_start = j[2]; kv_len = j[3]; mask_row = j[5];
    };

    if (warp == 0) {
        // ---------------------------------------------------------------- Q producer
        if (ptx::elect_one()) {
            int n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                int q_start, q_len, kv_start, kv_len, mask_row, head;
                item_of(item, q_start, q_len, kv_start, kv_len, mask_row, head);
                const int n_tiles = (q_len + TQ - 1) / TQ;
                for (int t = 0; t < n_tiles; ++t, ++n) {
                    const int b = n & 1;
                    mbar_wait_relaxed(&bars[XQ_FREE + b], ((n >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&bars[XQ_FULL + b], 2 * XQATOM);
                    ptx::tma_load_2d(smem + XQ_OFF + (2 * b) * XQATOM, &tmQ, &bars[XQ_FULL + b], head * XD, q_start + t * TQ);
                    ptx::tma_load_2d(smem + XQ_OFF + (2 * b + 1) * XQATOM, &tmQ, &bars[XQ_FULL + b], head * XD + 64, q_start + t * TQ);
                }
            }
        }
    } else if (warp == 3) {
        // ---------------------------------------------------------------- image K / V producer (double buffered per item)
        if (ptx::elect_one()) {
            uint32_t k = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
                int q_start, q_len, kv_start, kv_len, mask_row, head;
                item_of(item, q_start, q_len, kv_start, kv_len, mask_row, head);
                const int kb = k & 1;
                mbar_wait_relaxed(&bars[XKV_FREE + kb], ((k >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&bars[XKV_FULL + kb], 4 * XKATOM);
                ptx::tma_load_2d(smem + XKI_OFF + (2 * kb) * XKATOM, &tmK, &bars[XKV_FULL + kb], head * XD, kv_start);
                ptx::tma_load_2d(smem + XKI_OFF + (2 * kb + 1) * XKATOM, &tmK, &bars[XKV_FULL + kb], head * XD + 64, kv_start);
                ptx::tma_load_2d(smem + XVI_OFF + (2 * kb) * XKATOM, &tmV, &bars[XKV_FULL + kb], head * XD, kv_start);
                ptx::tma_load_2d(smem + XVI_OFF + (2 * kb + 1) * XKATOM, &tmV, &bars[XKV_FULL + kb], head * XD + 64, kv_start);
            }
            for (uint32_t j = (k > 2 ? k - 2 : 0); j < k; ++j)     // the last commits must have landed before this CTA exits
                mbar_wait_relaxed(&bars[XKV_FREE + (j & 1)], (j >> 1) & 1);
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer: QK(n + 1) ahead of PV(n)
        if (ptx::elect_one()) {
            const uint32_t fmt = FP16 ? 0u : 1u;
            struct Cursor { int item; uint32_t seq; int t, n, n_tiles, kv_len; bool valid; };
            auto load = [&](Cursor& c) {
                c.valid = c.item < n_items;
                if (c.valid) {
                    int q_start, q_len, kv_start, mask_row, head;
                    item_of(c.item, q_start, q_len, kv_start, c.kv_len, mask_row, head);
                    c.n_tiles = (q_len + TQ - 1) / TQ;
                }
            };
            auto start = [&](Cursor& c) { c.item = blockIdx.x; c.seq = 0; c.t = 0; c.n = 0; load(c); };
            auto advance = [&](Cursor& c) {
                ++c.n;
                if (++c.t == c.n_tiles) { c.t = 0; ++c.seq; c.item += gridDim.x; load(c); }
            };
            auto issue_qk = [&](const Cursor& c) {
                const int b = c.n & 1, kb = c.seq & 1;
                const uint32_t ph = (c.n >> 1) & 1;
                if (c.t == 0) ptx::mbar_wait(&bars[XKV_FULL + kb], (c.seq >> 1) & 1);
                ptx::mbar_wait(&bars[XQ_FULL + b], ph);
                ptx::mbar_wait(&bars[XS_FREE + b], ph ^ 1);
                ptx::tc_fence_after();
                const uint32_t idesc = ptx::make_idesc_f16(TQ, 64, fmt);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t da = ptx::make_sw128_kmajor_desc(smem_base + XQ_OFF + (2 * b + (ks >> 2)) * XQATOM) + 2 * (ks & 3);
                    const uint64_t db = ptx::make_sw128_kmajor_desc(smem_base + XKI_OFF + (2 * kb + (ks >> 2)) * XKATOM) + 2 * (ks & 3);
                    ptx::umma_f16_ss(tmem_base + b * 256, da, db, idesc, ks != 0 ? 1u : 0u);
                }
                ptx::umma_commit(&bars[XS_FULL + b]);
                ptx::umma_commit(&bars[XQ_FREE + b]);      // Q is read by these MMAs only
            };
            auto issue_pv = [&](const Cursor& c) {
                const int b = c.n & 1, kb = c.seq & 1;
                const uint32_t ph = (c.n >> 1) & 1;
                ptx::mbar_wait(&bars[XP_FULL + b], ph);
                ptx::tc_fence_after();
                const uint32_t idesc = ptx::make_idesc_f16(TQ, XD, fmt) | (1u << 16);      // B = V as stored: MN-major, two 64-dim atoms
                const uint64_t db = make_sw128_desc_lbo(smem_base + XVI_OFF + (2 * kb) * XKATOM, XKATOM);
                const int nks = (c.kv_len + 15) >> 4;                                      // 16 keys per MMA
                for (int ks = 0; ks < nks; ++ks)
                    ptx::umma_f16_ts(tmem_base + b * 256 + XO_COL, tmem_base + b * 256 + 8 * ks, db + 128 * ks, idesc, ks != 0 ? 1u : 0u);
                ptx::umma_commit(&bars[XO_FULL + b]);
                if (c.t == c.n_tiles - 1) ptx::umma_commit(&bars[XKV_FREE + kb]);
            };
            Cursor cq, cp;
            start(cq);
            start(cp);
            if (cq.valid) { issue_qk(cq); advance(cq); }
            while (cp.valid) {
                if (cq.valid) { issue_qk(cq); advance(cq); }
                issue_pv(cp);
                advance(cp);
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- softmax / epilogue warpgroups (thread = query row)
        const int wg = (warp - 4) >> 2;
        const int qd = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
        const float sl = a.scale * 1.4426950408889634f;
        const uint32_t ost = smem_base + XOST_OFF + (warp - 4) * 8192;
        int n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int q_start, q_len, kv_start, kv_len, mask_row, head;
            item_of(item, q_start, q_len, kv_start, kv_len, mask_row, head);
            const int n_tiles = (q_len + TQ - 1) / TQ;
            const int q_end = q_start + q_len;
            // key validity bits (image padding mask of the unit); no valid key = every key allowed, as the reference's additive mask
            unsigned long long kbits = kv_len >= 64 ? ~0ull : ((1ull << kv_len) - 1ull);
            if (mask_row >= 0) {
                const float* km = a.key_mask + static_cast<size_t>(mask_row) * a.key_mask_ld;
                const unsigned lo = __ballot_sync(0xffffffffu, lane < kv_len && km[lane] > 0.5f);
                const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < kv_len && km[min(lane + 32, kv_len - 1)] > 0.5f);
                const unsigned long long m = (static_cast<unsigned long long>(hi) << 32) | lo;
                if (m != 0ull) kbits = m;
            }
            bf16* Og = static_cast<bf16*>(a.o) + head * XD;
            for (int t = 0; t < n_tiles; ++t, ++n) {
                if ((n & 1) != wg) continue;
                const int b = wg;
                const uint32_t ph = (n >> 1) & 1;
                const int r0 = q_start + t * TQ;
                const uint32_t tS = tmem_base + b * 256 + lane_off;
                ptx::mbar_wait(&bars[XS_FULL + b], ph);
                ptx::tc_fence_after();
                uint32_t v[64];
                ptx::tmem_ld_32x32b_x32(tS, v);
                ptx::tmem_ld_32x32b_x32(tS + 32, v + 32);
                ptx::tmem_ld_wait();
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const float sv = ((kbits >> j) & 1ull) ? __uint_as_float(v[j]) : -INFINITY;
                    v[j] = __float_as_uint(sv);
                    mx = fmaxf(mx, sv);
                }
                const float msl = mx * sl;
                float l0 = 0.f, l1 = 0.f;
                uint32_t pk[32];
#pragma unroll
                for (int j = 0; j < 64; j += 2) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(v[j]), sl, -msl));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(v[j + 1]), sl, -msl));
                    l0 += p0; l1 += p1;
                    pk[j >> 1] = pack2<FP16>(p0, p1);
                }
                ptx::tmem_st_32x32b_x32(tS, pk);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars[XP_FULL + b]);
                const float inv = 1.0f / (l0 + l1);

                ptx::mbar_wait(&bars[XO_FULL + b], ph);
                ptx::tc_fence_after();
                // normalised row -> this warp's staging block: 32 rows x 256 B, 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t u[32];
                    ptx::tmem_ld_32x32b_x32(tS + XO_COL + c * 32, u);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int chunk = c * 4 + q;
                        ptx::sts_v4(ost + lane * 256 + ((chunk ^ (lane & 7)) << 4),
                                    pack2<FP16>(__uint_as_float(u[8 * q]) * inv, __uint_as_float(u[8 * q + 1]) * inv),
                                    pack2<FP16>(__uint_as_float(u[8 * q + 2]) * inv, __uint_as_float(u[8 * q + 3]) * inv),
                                    pack2<FP16>(__uint_as_float(u[8 * q + 4]) * inv, __uint_as_float(u[8 * q + 5]) * inv),
                                    pack2<FP16>(__uint_as_float(u[8 * q + 6]) * inv, __uint_as_float(u[8 * q + 7]) * inv));
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&bars[XS_FREE + b]);
                // whole 256-byte rows to global: 16 lanes per row, 2 rows per instruction
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int lr = i * 2 + (lane >> 4), ch = lane & 15;
                    const int row = r0 + qd * 32 + lr;
                    uint32_t x0, x1, x2, x3;
                    ptx::lds_v4(ost + lr * 256 + ((ch ^ (lr & 7)) << 4), x0, x1, x2, x3);
                    if (row < q_end) *reinterpret_cast<uint4*>(Og + static_cast<size_t>(row) * a.ldo + 8 * ch) = make_uint4(x0, x1, x2, x3);
                }
                __syncwarp();
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

template <bool FP16>
int launch_cross(const AttnJobsArgs& a, cudaStream_t stream) {
    CUtensorMap tmQ, tmK, tmV;
    const int W = a.heads * x::XD;
    UNIMM_TRY(gemm_make_map(a.q, a.n_rows, W, a.ldq, TQ, &tmQ));
    UNIMM_TRY(gemm_make_map(a.k, a.n_kv_rows, W, a.ldk, 64, &tmK));
    UNIMM_TRY(gemm_make_map(a.v, a.n_kv_rows, W, a.ldv, 64, &tmV));
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&attn_cross_umma_kernel<FP16>), x::XSMEM_BYTES));
    const int n_items = a.n_jobs * a.heads;
    const int grid = n_items < gemm_num_sms() ? n_items : gemm_num_sms();
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = x::XSMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    add_pdl_attr(attr, &cfg.numAttrs);
    UNIMM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_cross_umma_kernel<FP16>, tmQ, tmK, tmV, a, n_items));
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace

bool attention_candidates_umma_supported(const AttnJobsArgs& a, int halo) {
    return a.D == 64 && halo <= HALO && a.kv_cap <= 256 && a.n_rows > 0 && (a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 &&
           (a.ldo % 2) == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.k) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(a.v) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0 && (a.ldo % 8) == 0;
}

// candidate jobs (win = 1, D = 64, 16-bit, halo <= 16) on tcgen05: see the header of this file
int attention_candidates_umma(const AttnJobsArgs& a, int halo, cudaStream_t stream) {
    UNIMM_CHECK(a.n_jobs > 0 && attention_candidates_umma_supported(a, halo), "tcgen05 candidate attention: unsupported arguments");
    return a.lp_kind == LP_FP16 ? launch_umma<true, false>(a, stream) : launch_umma<false, false>(a, stream);
}

bool attention_dense_umma_supported(const AttnJobsArgs& a) {
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return a.D == 64 && a.seq_len > 0 && a.seq_len <= 256 && a.desc != nullptr && a.n_rows == a.n_jobs * a.seq_len && (a.ldq % 8) == 0 &&
           (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 8) == 0 && a16(a.q) && a16(a.k) && a16(a.v) && a16(a.o);
}
// text self-attention of the dense [B, S] layout under the descriptor masks on tcgen05 (n_jobs = B sequences, jobs = nullptr)
int attention_dense_umma(const AttnJobsArgs& a_in, cudaStream_t stream) {
    AttnJobsArgs a = a_in;
    a.jobs = nullptr;
    UNIMM_CHECK(a.n_jobs > 0 && attention_dense_umma_supported(a), "tcgen05 dense attention: unsupported arguments");
    return a.lp_kind == LP_FP16 ? launch_umma<true, true>(a, stream) : launch_umma<false, true>(a, stream);
}

}  // namespace unimm

namespace unimm {
bool attention_cross_umma_supported(const AttnJobsArgs& a) {
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return a.D == 128 && a.kv_cap <= 64 && a.win_cap == 0 && a.n_rows > 0 && a.n_kv_rows > 0 && (a.ldq % 8) == 0 && (a.ldk % 8) == 0 &&
           (a.ldv % 8) == 0 && (a.ldo % 8) == 0 && a16(a.q) && a16(a.k) && a16(a.v) && a16(a.o);
}
// jobs without windows over <= 64 keys, D = 128, 16-bit (text -> image co-attention) on tcgen05
int attention_cross_umma(const AttnJobsArgs& a, cudaStream_t stream) {
    UNIMM_CHECK(a.n_jobs > 0 && attention_cross_umma_supported(a), "tcgen05 cross attention: unsupported arguments");
    return a.lp_kind == LP_FP16 ? launch_cross<true>(a, stream) : launch_cross<false>(a, stream);
}
}  // namespace unimm
