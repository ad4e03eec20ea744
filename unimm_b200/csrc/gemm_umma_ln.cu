// tcgen05 GEMM with the residual add and the LayerNorm fused into the epilogue, one thread-block cluster per
// 128-row block:      X[M,N] = LayerNorm(A[M,K] · W[N,K]^T + bias + R[M,N]) * gamma + beta,   N = CS * 256
//
// Replaces, per sub-layer of the reference, nn.Linear -> dropout(eval: identity) -> "+ input_tensor" -> BertLayerNorm
// (models/vilbert_dialog.py:422-426 BertSelfOutput, :465-469 BertOutput, :548-554 / :592-597 image stream,
// :745-752 BertBiOutput, :1488-1492 image embeddings) — previously a GEMM that wrote the fp32 pre-LN tensor and a
// LayerNorm kernel that read it back (18 B of HBM traffic per element and a second launch; now 10 B, one launch).
//
// A LayerNorm row is 768 or 1024 fp32 accumulator columns, more than the 512 TMEM columns one CTA owns, so the row
// is split over a cluster of CS = N/256 CTAs (3 or 4): CTA `rank` computes the 128 x 256 tile of columns
// [256*rank, 256*rank+256) of the cluster's row block with the same warp-specialised TMA -> tcgen05.mma pipeline as
// gemm_umma.cu, and the epilogue exchanges per-row (sum, M2) partials through distributed shared memory:
//
//   precharge  before the MMA warp may touch accumulator buffer g for a tile, the epilogue group that owns g loads the
//              tile's residual rows (+ bias) and writes them INTO the accumulator with tcgen05.st; every tcgen05.mma of
//              the tile then accumulates, so "A W^T + bias + residual" leaves the tensor core complete.  The loads do
//              not depend on anything, so three 32-column chunks per thread are kept in flight (and the rows of the
//              tile after that are prefetched into L2): the HBM latency of the residual stream is paid off the critical
//              path, between two tiles, instead of once per chunk inside the epilogue.
//   pass 1     tcgen05.ld (16x256b fragments) -> per-row sum / sum of squares of this CTA's 256 columns
//   exchange   each warp pushes its 32 rows' (sum, M2 about the local mean) into every peer CTA's shared memory with
//              st.async (data + mbarrier complete_tx in one async-proxy operation, no cluster-scope fences); Chan's
//              formula merges the CS partials, so the variance never suffers the E[x^2] - E[x]^2 cancellation across tiles
//   pass 2     tcgen05.ld again -> (v - mean) * rstd * gamma + beta -> fp32 master + 16-bit shadow stores
//
// No shared-memory transpose: in the 16x256b fragment a thread owns accumulator columns {8 j + 2 a + e} (a = lane % 4,
// j = 0..3, e = 0..1) of a 32-column chunk.  The weight rows of every 32-row group are stored pre-permuted
// (ln_weight_row(), applied once at load time by permute_weight_rows_ln) so that those 8 accumulator columns ARE output
// columns {4 a .. 4 a + 3} and {16 + 4 a .. 16 + 4 a + 3}: each 128-bit global access of four neighbouring lanes
// covers 64 contiguous bytes (whole 32-byte sectors), and TMEM doubles as the only staging buffer.
//
// The 8 epilogue warps form two groups of 4; group g owns accumulator buffer g and the tiles of parity g, so the
// exchange latency and the precharge of one buffer overlap the other group's passes and the other tile's main loop.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace unimm {

namespace {

constexpr int BM = 128;
constexpr int IDN = 256;           // the identity operand (largest column block per CTA)
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr float kLnEps = 1e-12f;   // BertLayerNorm eps (reference models/vilbert_dialog.py:322)

template <int CS, int BN, bool PAIR = false>
struct LnCfg {
    static constexpr int kStages = PAIR ? 6 : (BN == 256 ? 4 : 5);
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * BK * 2;           // a cta_group::2 pair keeps half of the W box per CTA
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kIdentBytes = (PAIR ? 32 : 64) * BK * 2;           // resident 64 x 64 identity block (a pair keeps half of its rows per CTA)
    static constexpr int kStatsBytes = 2 * 2 * 4 * (CS - 1) * 32 * 8;       // [use parity][group][warp][source][row] float2
    static constexpr int kParamBytes = 3 * BN * 4;                          // bias, gamma, beta of this CTA's 256 columns
    static constexpr int kBarBytes = 512;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kIdentBytes + kStatsBytes + kParamBytes + kBarBytes;
};

// RES16 = false: fp32 residual, precharged into the accumulator by the epilogue warps (see above).
// RES16 = true : the residual is the 16-bit activation copy itself (fp16 residual stream).  A 16-bit residual is an
//                exact tensor-core operand, so it is added by the tensor core: four extra k-blocks per tile multiply
//                the residual tile R[128 x 256] (TMA box from tmR) by a (row-permuted) 256 x 256 identity (tmI) and
//                accumulate in fp32 — no register traffic, no latency-exposed loads, the epilogue never touches global
//                memory except for its stores.  Only the 64 x 64 diagonal block of each identity k-block is non-zero (the row
//                permutation is 32-periodic), so residual k-block j is issued as N = 64 MMAs into accumulator columns
//                [64 j, 64 j + 64) against ONE 64 x 64 identity block that stays resident in shared memory (loaded once per CTA):
//                +8 % MMA work at K = 768 (+2 % at 3072) instead of +33 % (+8 %), and no identity operand traffic at all.
// BN = columns per CTA: 256 (N = 768 as 3 CTAs, N = 1024 as 4) or 192 (N = 768 as 4 CTAs: clusters of 4 tile all 148 SMs, clusters
// of 3 only 135 of them).
// RLP (with RES16 = false): the precharged residual is the 16-bit activation copy (converted by the epilogue warps) — the
// residual costs no tensor-core work at all, only its 2 bytes per element of HBM read.
// PAIR: clusters of 2 * CS CTAs — two vertically adjacent row blocks per cluster, the two CTAs that own the same 256 columns of
// them form a cta_group::2 pair (ranks 2 n and 2 n + 1): ONE M = 256 MMA issued by the even rank, half of the W box (and of the
// identity) in each CTA, TMA completions on the leader's barrier, multicast commits (see gemm_umma.cu, CL = 3).  The row
// statistics are exchanged among the CS CTAs of the same row block (ranks of equal parity).
template <int CS, bool RES16, int BN, bool RLP = false, bool PAIR = false>
__global__ void __launch_bounds__(384, 1)
umma_gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmI, int M, int K, GemmLnEpilogue ep) {
    using Cfg = LnCfg<CS, BN, PAIR>;
    constexpr int N = CS * BN;
    constexpr int CL = PAIR ? 2 * CS : CS;                   // CTAs per cluster
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::kStages * Cfg::kABytes;
    uint8_t* sI = smem + Cfg::kStages * Cfg::kStageBytes;    // 1024-byte aligned: the stage sizes are multiples of 1024
    float2* stats = reinterpret_cast<float2*>(sI + Cfg::kIdentBytes);
    float* sparam = reinterpret_cast<float*>(sI + Cfg::kIdentBytes + Cfg::kStatsBytes);   // [3][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sI + Cfg::kIdentBytes + Cfg::kStatsBytes + Cfg::kParamBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* stats_bar = tempty_bar + 2;                  // [use parity][group][warp] = 16 barriers
    uint64_t* ident_bar = stats_bar + 16;                  // the resident identity block has landed (once per kernel)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ident_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = ptx::cluster_ctarank();
    const uint32_t rank = PAIR ? crank >> 1 : crank;         // column block of this CTA
    const int m_r = PAIR ? static_cast<int>(crank & 1) : 0;  // which row block of the pair
    const uint32_t leader = crank & ~1u;                     // PAIR: the rank that issues the MMAs for this CTA
    const uint16_t pair_mask = static_cast<uint16_t>(3u << leader);
    const int cluster_id = blockIdx.x / CL;
    const int num_clusters = gridDim.x / CL;
    const int num_m = PAIR ? ((M + BM - 1) / BM + 1) / 2 : (M + BM - 1) / BM;      // row blocks (pairs of them) = work items of a cluster
    auto row_of = [&](int mb) { return (PAIR ? 2 * mb + m_r : mb) * BM; };          // first row of this CTA's block in work item mb
    const int num_k = K / BK;
    const int num_kr = RES16 ? num_k + BN / BK : num_k;       // + the residual x identity k-blocks
    const int n0 = static_cast<int>(rank) * BN;

    if (warp == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && ptx::elect_one()) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], (ep.a_multicast && !PAIR) ? CS : 1);   // multicast A: a slot is reusable once EVERY CTA of the cluster has read it
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], PAIR ? 8 : 4);   // one arrival per epilogue warp of the group (of both CTAs of a pair)
        }
        for (int b = 0; b < 16; ++b) ptx::mbar_init(&stats_bar[b], 1);   // the owner's arrive.expect_tx; peers complete bytes
        ptx::mbar_init(ident_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) ptx::tmem_alloc_2sm<512>(tmem_slot);
        else ptx::tmem_alloc<512>(tmem_slot);
    }
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
        sparam[i] = ep.bias[n0 + i];
        sparam[BN + i] = ep.gamma[n0 + i];
        sparam[2 * BN + i] = ep.beta[n0 + i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();    // every peer's barriers are initialised before anyone arrives on them remotely
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch_dependents();
    ptx::grid_dep_wait();                  // everything above (weights-only reads included) overlapped the previous kernel's tail

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            if (RES16) {      // the 64 x 64 identity block (rows [32 m_r, 32 m_r + 32) of it in a pair): once, stays resident
                if (PAIR) {
                    if (m_r == 0) ptx::mbar_arrive_expect_tx(ident_bar, 2 * Cfg::kIdentBytes);
                    ptx::tma_load_2d_2sm(sI, &tmI, ident_bar, 0, m_r * 32);
                } else {
                    ptx::mbar_arrive_expect_tx(ident_bar, Cfg::kIdentBytes);
                    ptx::tma_load_2d(sI, &tmI, ident_bar, 0, 0);
                }
            }
            for (int mb = cluster_id; mb < num_m; mb += num_clusters) {
                const int m0 = row_of(mb);
                for (int kb = 0; kb < num_kr; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    const bool resid = RES16 && kb >= num_k;      // residual k-block: only the A box (the residual tile) is loaded
                    if (PAIR) {
                        if (m_r == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], resid ? 2 * Cfg::kABytes : 2 * Cfg::kStageBytes);     // both CTAs' boxes
                        if (!resid) {
                            ptx::tma_load_2d_2sm(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * BK, m0);
                            ptx::tma_load_2d_2sm(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, n0 + m_r * (BN / 2));
                        } else {
                            const int j = kb - num_k;
                            ptx::tma_load_2d_2sm(sA + stage * Cfg::kABytes, &tmR, &full_bar[stage], n0 + j * BK, m0);
                        }
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], resid ? Cfg::kABytes : Cfg::kStageBytes);
                    if (!resid) {
                        // the CS CTAs of the cluster multiply the SAME 128 x 64 activation box: CTA kb % CS fetches it once and
                        // multicasts it into every CTA's slot (data and mbarrier bytes land at the same CTA-relative offsets)
                        if (!ep.a_multicast) ptx::tma_load_2d(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * BK, m0);
                        else if (kb % CS == static_cast<int>(rank))   /* (a_multicast is never set with PAIR) */
                            ptx::tma_load_2d_mc(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * BK, m0, static_cast<uint16_t>((1u << CS) - 1u));
                        ptx::tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * BK, n0);
                    } else {      // residual columns n0 + 64 j .. (against the resident identity block)
                        const int j = kb - num_k;
                        ptx::tma_load_2d(sA + stage * Cfg::kABytes, &tmR, &full_bar[stage], n0 + j * BK, m0);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (ptx::elect_one() && m_r == 0) {
            const uint32_t in_fmt = (ep.in_kind < 0 ? ep.lp_kind : ep.in_kind) == LP_FP16 ? 0u : 1u;
            const uint32_t idesc = ptx::make_idesc_f16(PAIR ? 2 * BM : BM, BN, in_fmt);
            // the residual x identity k-blocks are instructions of their own: they read the RESIDUAL's format (out_lp's), which may differ
            // from the A / W format of the projection itself (bf16 mode: fp16 LayerNorm outputs, bf16 projection operands) — tcgen05
            // only wants one format per instruction, and both kinds accumulate into the same fp32 TMEM columns
            const uint32_t idesc_r = ptx::make_idesc_f16(PAIR ? 2 * BM : BM, BK, ep.lp_kind == LP_FP16 ? 0u : 1u);   // N = 64: one diagonal block
            const uint64_t di = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sI));
            if (RES16) {
                ptx::mbar_wait(ident_bar, 0);
                ptx::tc_fence_after();
            }
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int mb = cluster_id; mb < num_m; mb += num_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                // RES16: buffer drained by its group.  fp32 residual: the group has precharged it with residual + bias.
                ptx::mbar_wait(&tempty_bar[acc], RES16 ? (acc_phase ^ 1) : acc_phase);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kr; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + stage * Cfg::kABytes));
                    if (RES16 && kb >= num_k) {
                        // accumulator columns [64 j, 64 j + 64) += residual columns [64 j, 64 j + 64) x the (row-permuted) identity block
                        const uint32_t tmem_r = tmem_d + static_cast<uint32_t>(kb - num_k) * BK;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            if (PAIR) ptx::umma_f16_ss_2sm(tmem_r, da + 2 * k, di + 2 * k, idesc_r, 1u);
                            else ptx::umma_f16_ss(tmem_r, da + 2 * k, di + 2 * k, idesc_r, 1u);
                        }
                    } else {
                        const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB + stage * Cfg::kBBytes));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            if (PAIR) ptx::umma_f16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (RES16 && (kb | k) == 0) ? 0u : 1u);
                            else ptx::umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (RES16 && (kb | k) == 0) ? 0u : 1u);
                        }
                    }
                    if (PAIR) ptx::umma_commit_2sm_mc(&empty_bar[stage], pair_mask);
                    else if (ep.a_multicast) ptx::umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>((1u << CS) - 1u));
                    else ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (PAIR) ptx::umma_commit_2sm_mc(&tfull_bar[acc], pair_mask);
                else ptx::umma_commit(&tfull_bar[acc]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue: group g = tiles of parity g
        const int g = (warp - 4) >> 2;
        const int q = (warp - 4) & 3;            // TMEM lane quarter (= warp % 4)
        const int a = lane & 3;                  // which 2 x 4 columns of a 32-column chunk this lane owns
        const int r8 = lane >> 2;                // rows 8 k + r8, k = 0..3, of the warp's 32 rows
        const uint32_t taddr0 = tmem_base + g * BN + (static_cast<uint32_t>(q * 32) << 16);
        constexpr int NCH = BN / 32;
        // one 32-row x 32-column chunk = two 16-lane slabs of 16 registers; register of (row 8 k + r8, value m):
        // m < 4 -> output column 4 a + m, m >= 4 -> 16 + 4 a + (m - 4) of the chunk
        // the accumulator goes back to the MMA warp of the pair's LEADER
        auto arrive_tempty = [&](uint64_t* bar) {
            if (PAIR && m_r != 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), leader));
            else ptx::mbar_arrive(bar);
        };
        auto frag = [](int k, int m) { return (k >> 1) * 16 + 4 * (m >> 1) + 2 * (k & 1) + (m & 1); };
        auto ld_chunk = [&](int c, uint32_t* v) {
            ptx::tmem_ld_16x256b_x4(taddr0 + c * 32, v);
            ptx::tmem_ld_16x256b_x4(taddr0 + c * 32 + (16u << 16), v + 16);
        };
        const float* sbias = sparam;
        const float* sgamma = sparam + BN;
        const float* sbeta = sparam + 2 * BN;

        // accumulator <- residual + bias for row block mb (tcgen05.st), then hand the buffer to the MMA warp
        auto precharge = [&](int mb) {
            const int row0 = row_of(mb) + q * 32 + r8;
            const float* res_base = RLP ? nullptr : ep.residual + static_cast<size_t>(row0) * ep.ldr + n0 + a * 4;
            const bf16* res_lp = RLP ? ep.residual_lp + static_cast<size_t>(row0) * ep.ldr_lp + n0 + a * 4 : nullptr;
            auto cvt4 = [&](uint2 w) {
                float2 lo, hi;
                if (ep.lp_kind == LP_FP16) {
                    lo = __half22float2(*reinterpret_cast<const __half2*>(&w.x)); hi = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
                } else {
                    lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x)); hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
                }
                return make_float4(lo.x, lo.y, hi.x, hi.y);
            };
            auto ld_res = [&](int c, float4* r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (row0 + 8 * k < M) {
                        if (RLP) {
                            const bf16* p = res_lp + static_cast<size_t>(8 * k) * ep.ldr_lp + c * 32;
                            r[2 * k] = cvt4(*reinterpret_cast<const uint2*>(p));
                            r[2 * k + 1] = cvt4(*reinterpret_cast<const uint2*>(p + 16));
                            continue;
                        }
                        const float* p = res_base + static_cast<size_t>(8 * k) * ep.ldr + c * 32;
                        r[2 * k] = *reinterpret_cast<const float4*>(p);
                        r[2 * k + 1] = *reinterpret_cast<const float4*>(p + 16);
                    } else {
                        r[2 * k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        r[2 * k + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            };
            {   // pull the rows of this group's tile AFTER this one towards L2
                const int mb2 = mb + 2 * num_clusters;
                const int prow = row_of(mb2) + q * 32 + lane;
                if (mb2 < num_m && prow < M) {
                    const char* p = RLP ? reinterpret_cast<const char*>(ep.residual_lp + static_cast<size_t>(prow) * ep.ldr_lp + n0)
                                        : reinterpret_cast<const char*>(ep.residual + static_cast<size_t>(prow) * ep.ldr + n0);
#pragma unroll
                    for (int i = 0; i < BN * (RLP ? 2 : 4) / 128; ++i) ptx::prefetch_l2(p + i * 128);
                }
            }
            float4 r[3][8];
            ld_res(0, r[0]);
            ld_res(1, r[1]);
            ld_res(2, r[2]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float4 b0 = *reinterpret_cast<const float4*>(sbias + c * 32 + a * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(sbias + c * 32 + 16 + a * 4);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {            // one 16-lane slab (rows 8 k + r8, k = 2 sl, 2 sl + 1) at a time
                    uint32_t w[16];
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int k = 2 * sl + kk;
                        const float4 x0 = r[c % 3][2 * k], x1 = r[c % 3][2 * k + 1];
                        w[frag(kk, 0)] = __float_as_uint(x0.x + b0.x); w[frag(kk, 1)] = __float_as_uint(x0.y + b0.y);
                        w[frag(kk, 2)] = __float_as_uint(x0.z + b0.z); w[frag(kk, 3)] = __float_as_uint(x0.w + b0.w);
                        w[frag(kk, 4)] = __float_as_uint(x1.x + b1.x); w[frag(kk, 5)] = __float_as_uint(x1.y + b1.y);
                        w[frag(kk, 6)] = __float_as_uint(x1.z + b1.z); w[frag(kk, 7)] = __float_as_uint(x1.w + b1.w);
                    }
                    ptx::tmem_st_16x256b_x4(taddr0 + c * 32 + (static_cast<uint32_t>(16 * sl) << 16), w);
                }
                if (c + 3 < NCH) ld_res(c + 3, r[c % 3]);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(&tempty_bar[g]);
        };

        int mb = cluster_id + g * num_clusters;
        if (!RES16 && mb < num_m) precharge(mb);
        for (int use = 0; mb < num_m; mb += 2 * num_clusters, ++use) {
            const uint32_t par = use & 1;
            const int row0 = row_of(mb) + q * 32 + r8;            // this lane's rows: row0 + 8 k
            float2* slots = stats + ((par * 2 + g) * 4 + q) * (CS - 1) * 32;
            uint64_t* sbar = &stats_bar[(par * 2 + g) * 4 + q];
            if (lane == 0) ptx::mbar_arrive_expect_tx(sbar, (CS - 1) * 32 * 8);   // the peers' partials of this tile
            ptx::mbar_wait(&tfull_bar[g], par);
            ptx::tc_fence_after();

            // ---------------- pass 1: row statistics of v = A W^T + bias + residual over this CTA's 256 columns
            float s[4], ss[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { s[k] = 0.f; ss[k] = 0.f; }
            uint32_t v[2][32];
            ld_chunk(0, v[0]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                ptx::tmem_ld_wait();
                if (c + 1 < NCH) ld_chunk(c + 1, v[(c + 1) & 1]);
#pragma unroll
                float bb[8];
                if (RES16) {     // (the fp32-residual path has the bias in the accumulator already)
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + c * 32 + a * 4);
                    const float4 b1 = *reinterpret_cast<const float4*>(sbias + c * 32 + 16 + a * 4);
                    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        float y = __uint_as_float(v[c & 1][frag(k, m)]);
                        if (RES16) y += bb[m];
                        s[k] += y;
                        ss[k] = fmaf(y, y, ss[k]);
                    }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int o = 1; o < 4; o <<= 1) {
                    s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
                    ss[k] += __shfl_xor_sync(0xffffffffu, ss[k], o);
                }
            }
            float my_s = s[0], my_ss = ss[0];          // lane (r8, a) is responsible for row 8 a + r8
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (a == k) { my_s = s[k]; my_ss = ss[k]; }
            const float my_mean_c = my_s * (1.0f / BN);
            const float my_m2 = fmaxf(my_ss - my_s * my_mean_c, 0.f);
            const int my_row = 8 * a + r8;
#pragma unroll
            for (int p = 0; p < CS; ++p) {
                if (p == static_cast<int>(rank)) continue;
                const int src = static_cast<int>(rank) < p ? static_cast<int>(rank) : static_cast<int>(rank) - 1;
                const uint32_t pr = PAIR ? static_cast<uint32_t>(2 * p + m_r) : static_cast<uint32_t>(p);   // same row block, column block p
                ptx::st_async_v2(ptx::mapa(ptx::smem_u32(slots + src * 32 + my_row), pr), my_s, my_m2, ptx::mapa(ptx::smem_u32(sbar), pr));
            }
            ptx::mbar_wait(sbar, (use >> 1) & 1);
            float tot = my_s, m2 = my_m2;
            float pm[CS - 1];
#pragma unroll
            for (int j = 0; j < CS - 1; ++j) {
                const float2 t = slots[j * 32 + my_row];
                tot += t.x;
                m2 += t.y;
                pm[j] = t.x * (1.0f / BN);
            }
            const float mean = tot * (1.0f / N);
            float dev2 = (my_mean_c - mean) * (my_mean_c - mean);
#pragma unroll
            for (int j = 0; j < CS - 1; ++j) dev2 = fmaf(pm[j] - mean, pm[j] - mean, dev2);
            const float var = fmaf(dev2, static_cast<float>(BN), m2) * (1.0f / N);
            const float rstd = rsqrtf(var + kLnEps);
            float na[4], nb[4];     // normalised = v * na + nb
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float mi = __shfl_sync(0xffffffffu, mean, r8 * 4 + k);
                const float ri = __shfl_sync(0xffffffffu, rstd, r8 * 4 + k);
                na[k] = ri;
                nb[k] = -mi * ri;
            }

            // ---------------- pass 2: normalise and store
            ld_chunk(0, v[0]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int cl = c * 32 + a * 4;              // first of this lane's two 4-column pieces, within the tile
                const float4 g0 = *reinterpret_cast<const float4*>(sgamma + cl), g1 = *reinterpret_cast<const float4*>(sgamma + cl + 16);
                const float4 e0 = *reinterpret_cast<const float4*>(sbeta + cl), e1 = *reinterpret_cast<const float4*>(sbeta + cl + 16);
                const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                float bb[8];
                if (RES16) {
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + cl), b1 = *reinterpret_cast<const float4*>(sbias + cl + 16);
                    bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
                }
                ptx::tmem_ld_wait();
                if (c + 1 < NCH) ld_chunk(c + 1, v[(c + 1) & 1]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float o[8];
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        float y = __uint_as_float(v[c & 1][frag(k, m)]);
                        if (RES16) y += bb[m];
                        o[m] = fmaf(fmaf(y, na[k], nb[k]), gg[m], ee[m]);
                    }
                    const int grow = row0 + 8 * k;
                    if (grow < M) {
                        if (ep.out_f32 != nullptr) {
                            float* dst = ep.out_f32 + static_cast<size_t>(grow) * ep.ldo_f32 + n0 + cl;
                            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4*>(dst + 16) = make_float4(o[4], o[5], o[6], o[7]);
                        }
                        if (ep.out_lp != nullptr) {
                            bf16* dst = ep.out_lp + static_cast<size_t>(grow) * ep.ldo_lp + n0 + cl;
                            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_lp2(o[0], o[1], ep.lp_kind), pack_lp2(o[2], o[3], ep.lp_kind));
                            *reinterpret_cast<uint2*>(dst + 16) = make_uint2(pack_lp2(o[4], o[5], ep.lp_kind), pack_lp2(o[6], o[7], ep.lp_kind));
                        }
                    }
                }
            }
            if (RES16) {      // hand the drained accumulator back to the MMA warp
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_tempty(&tempty_bar[g]);
            } else if (mb + 2 * num_clusters < num_m) {
                precharge(mb + 2 * num_clusters);   // the buffer's next tile (which also hands it back to the MMA warp)
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();    // no CTA leaves while a peer may still push statistics into its shared memory
    if (warp == 2) {
        if (PAIR) ptx::tmem_dealloc_2sm<512>(tmem_base);
        else ptx::tmem_dealloc<512>(tmem_base);
    }
}

// weight row that must sit at position `p` of the B operand so that accumulator column p holds output column
// ln_weight_row(p): within each group of 32, p = 8 j + 2 a + e (the 16x256b fragment of lane a) holds output column
// 16 (j / 2) + 4 a + 2 (j % 2) + e, i.e. lane a owns output columns 4a..4a+3 and 16+4a..16+4a+3
__host__ __device__ inline int ln_weight_row(int p) {
    const int j = (p >> 3) & 3, a = (p >> 1) & 3, e = p & 1;
    return (p & ~31) | (16 * (j >> 1) + 4 * a + 2 * (j & 1) + e);
}

// the 16-bit-output epilogue of gemm_umma.cu: accumulator column 8 j + 2 a + e holds output column 8 a + 2 j + e,
// i.e. lane a owns the 8 consecutive output columns 8a..8a+7 of a 32-column chunk
__host__ __device__ inline int p16_weight_row(int p) {
    const int j = (p >> 3) & 3, a = (p >> 1) & 3, e = p & 1;
    return (p & ~31) | (8 * a + 2 * j + e);
}

__global__ void permute_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int row_vec, int mode) {
    const int row = blockIdx.x;
    const uint4* s = src + static_cast<size_t>(mode == 0 ? ln_weight_row(row) : p16_weight_row(row)) * row_vec;
    uint4* d = dst + static_cast<size_t>(row) * row_vec;
    for (int i = threadIdx.x; i < row_vec; i += blockDim.x) d[i] = s[i];
}

// 256 x 256 identity in the 16-bit encoding `lp_kind`, rows permuted like the weights (accumulator column p receives
// residual column ln_weight_row(p)); built once per device and encoding
__global__ void identity_kernel(uint16_t* I, uint16_t one) {
    const int p = blockIdx.x;
    for (int k = threadIdx.x; k < IDN; k += blockDim.x) I[p * IDN + k] = (k == ln_weight_row(p)) ? one : uint16_t(0);
}
int identity_for(int lp_kind, const bf16** out) {
    static const bf16* cache[16][2] = {};
    int dev = 0;
    UNIMM_CUDA_CHECK(cudaGetDevice(&dev));
    UNIMM_CHECK(dev >= 0 && dev < 16, "device index out of range");
    const int kind = lp_kind == LP_FP16 ? 1 : 0;
    if (cache[dev][kind] == nullptr) {
        void* p = nullptr;
        UNIMM_CUDA_CHECK(cudaMalloc(&p, IDN * IDN * 2));
        identity_kernel<<<IDN, 128>>>(static_cast<uint16_t*>(p), kind ? uint16_t(0x3C00) : uint16_t(0x3F80));
        UNIMM_CUDA_CHECK(cudaDeviceSynchronize());
        cache[dev][kind] = static_cast<const bf16*>(p);
    }
    *out = cache[dev][kind];
    return 0;
}

template <int CS, bool RES16, int BN, bool RLP = false, bool PAIR = false>
int launch_ln(const bf16* A, int lda, const bf16* W, int ldw, int M, int K, const GemmLnEpilogue& ep, cudaStream_t stream) {
    using Cfg = LnCfg<CS, BN, PAIR>;
    constexpr int CL = PAIR ? 2 * CS : CS;
    CUtensorMap tmA, tmB, tmR, tmI;
    UNIMM_TRY(gemm_make_map(A, M, K, lda, BM, &tmA));
    UNIMM_TRY(gemm_make_map(W, CS * BN, K, ldw, PAIR ? BN / 2 : BN, &tmB));
    if (RES16) {
        const bf16* ident = nullptr;
        UNIMM_TRY(identity_for(ep.lp_kind, &ident));
        UNIMM_TRY(gemm_make_map(ep.residual_lp, M, CS * BN, ep.ldr_lp, BM, &tmR));
        UNIMM_TRY(gemm_make_map(ident, IDN, IDN, IDN, PAIR ? 32 : 64, &tmI));          // top-left 64 x 64 block: the row order is 32-periodic
    } else {
        tmR = tmA;
        tmI = tmB;
    }
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&umma_gemm_ln_kernel<CS, RES16, BN, RLP, PAIR>), Cfg::kSmemBytes));
    if (max_clusters == 0) {
        cfg.gridDim = dim3((gemm_num_sms() / CL) * CL, 1, 1);
        int n = 0;
        UNIMM_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, umma_gemm_ln_kernel<CS, RES16, BN, RLP, PAIR>, &cfg));
        UNIMM_CHECK(n > 0, "no co-resident cluster fits the LayerNorm-fused GEMM");
        max_clusters = n < gemm_num_sms() / CL ? n : gemm_num_sms() / CL;
        if (getenv("UNIMM_DEBUG")) fprintf(stderr, "[unimm] LayerNorm-fused GEMM: cluster size %d, %d co-resident clusters (occupancy query %d)\n", CS, max_clusters, n);
    }
    const int num_m = PAIR ? ((M + BM - 1) / BM + 1) / 2 : (M + BM - 1) / BM;
    const int clusters = num_m < max_clusters ? num_m : max_clusters;
    cfg.gridDim = dim3(clusters * CL, 1, 1);
    add_pdl_attr(attr, &cfg.numAttrs);
    UNIMM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, umma_gemm_ln_kernel<CS, RES16, BN, RLP, PAIR>, tmA, tmB, tmR, tmI, M, K, ep));
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace

int permute_weight_rows(const bf16* W, bf16* Wp, int N, int K, int mode, cudaStream_t stream) {
    UNIMM_CHECK(N % 32 == 0 && K % 8 == 0 && W != Wp && (mode == 0 || mode == 1), "permute_weight_rows: N must be a multiple of 32, K of 8, out of place");
    permute_rows_kernel<<<N, 128, 0, stream>>>(reinterpret_cast<const uint4*>(W), reinterpret_cast<uint4*>(Wp), N, K / 8, mode);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}
int permute_weight_rows_ln(const bf16* W, bf16* Wp, int N, int K, cudaStream_t stream) { return permute_weight_rows(W, Wp, N, K, 0, stream); }

bool gemm_umma_ln_supported(int N, int K, const GemmLnEpilogue& ep) {
    auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool res_ok = ep.residual_lp != nullptr ? (a16(ep.residual_lp) && (ep.ldr_lp & 7) == 0)
                                                  : (ep.residual != nullptr && a16(ep.residual) && (ep.ldr & 3) == 0);
    return (N == 768 || N == 1024) && K > 0 && K % BK == 0 && ep.bias && ep.gamma && ep.beta && a16(ep.bias) && res_ok &&
           (ep.out_f32 == nullptr || (a16(ep.out_f32) && (ep.ldo_f32 & 3) == 0)) &&
           (ep.out_lp == nullptr || ((reinterpret_cast<uintptr_t>(ep.out_lp) & 7) == 0 && (ep.ldo_lp & 3) == 0));
}

int gemm_umma_ln(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmLnEpilogue& ep_in, cudaStream_t stream) {
    UNIMM_CHECK(M > 0 && gemm_umma_ln_supported(N, K, ep_in), "LayerNorm-fused GEMM: N must be 768 or 1024, K a multiple of 64, pointers 16-byte aligned");
    // measured on B200 (profiles/r01_v8): sharing the activation box across the 3 CTAs by multicast couples their pipelines and
    // costs 2-4 % on FFN-2 + LN, more than the saved L2 traffic returns; kept as an opt-in (UNIMM_LN_MULTICAST=1)
    static const bool mc_enabled = getenv("UNIMM_LN_MULTICAST") != nullptr && atoi(getenv("UNIMM_LN_MULTICAST")) != 0;
    GemmLnEpilogue ep = ep_in;
    ep.a_multicast = ep.a_multicast && mc_enabled;
    // N = 768 as 4 CTAs x 192 columns keeps every SM busy (37 clusters of 4 = 148 SMs; 3 x 256 leaves 13 SMs idle) at the price of 11 %
    // more operand bytes per FLOP.  Measured (profiles/r01_v8): 6-10 % faster per launch in isolation, but 0.7 % SLOWER inside the
    // power-capped step (the step is limited by energy per FLOP, not by idle SMs) -> 3 x 256 stays the default, UNIMM_LN_SPLIT=4 opts in
    static const int split = getenv("UNIMM_LN_SPLIT") ? atoi(getenv("UNIMM_LN_SPLIT")) : 3;
    // 16-bit residual: added on the tensor core (identity k-blocks; UNIMM_LN_RES=mma) or precharged into the accumulator by the
    // epilogue warps (UNIMM_LN_RES=precharge: no extra MMA work, +25 % K otherwise at K = 768)
    static const bool res_precharge = getenv("UNIMM_LN_RES") != nullptr && std::string(getenv("UNIMM_LN_RES")) == "precharge";
    if (ep.residual_lp != nullptr && res_precharge) {
        if (N == 768) return split == 3 ? launch_ln<3, false, 256, true>(A, lda, W, ldw, M, K, ep, stream) : launch_ln<4, false, 192, true>(A, lda, W, ldw, M, K, ep, stream);
        return launch_ln<4, false, 256, true>(A, lda, W, ldw, M, K, ep, stream);
    }
    // tall problems with the 16-bit residual: cta_group::2 pairs inside clusters of 6 (N = 768) — UNIMM_LN_PAIR=0 keeps clusters of 3
    static const bool pair_enabled = getenv("UNIMM_LN_PAIR") == nullptr || atoi(getenv("UNIMM_LN_PAIR")) != 0;
    if (ep.residual_lp != nullptr && pair_enabled && N == 768 && M >= 8192 && split == 3) {
        ep.a_multicast = false;
        return launch_ln<3, true, 256, false, true>(A, lda, W, ldw, M, K, ep, stream);
    }
    if (ep.residual_lp != nullptr) {
        if (N == 768) return split == 3 ? launch_ln<3, true, 256>(A, lda, W, ldw, M, K, ep, stream) : launch_ln<4, true, 192>(A, lda, W, ldw, M, K, ep, stream);
        return launch_ln<4, true, 256>(A, lda, W, ldw, M, K, ep, stream);
    }
    // fp32 residual (the bf16 mode's stream): the same cta_group::2 pairs, the residual precharged by both CTAs' epilogue warps.
    // Same-box A/B of the bf16 step (profiles/r02_v12_ln_pair_f32_ab.txt): 298.3 k -> 305.2 k candidates / s; UNIMM_LN_PAIR_F32=0 opts out
    static const bool pair_f32 = getenv("UNIMM_LN_PAIR_F32") == nullptr || atoi(getenv("UNIMM_LN_PAIR_F32")) != 0;
    if (pair_f32 && N == 768 && M >= 8192 && split == 3) {
        ep.a_multicast = false;
        return launch_ln<3, false, 256, false, true>(A, lda, W, ldw, M, K, ep, stream);
    }
    if (N == 768) return split == 3 ? launch_ln<3, false, 256>(A, lda, W, ldw, M, K, ep, stream) : launch_ln<4, false, 192>(A, lda, W, ldw, M, K, ep, stream);
    return launch_ln<4, false, 256>(A, lda, W, ldw, M, K, ep, stream);
}

}  // namespace unimm
