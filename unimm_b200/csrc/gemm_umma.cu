// tcgen05 / TMEM / TMA GEMM for sm_100a:   C[M,N] = act(A[M,K] · W[N,K]^T + bias) (+ residual)
//
// Replaces the reference's nn.Linear call sites on the hot path (models/vilbert_dialog.py:386-388 QKV,
// :423 out-proj, :453 FFN-1, :466 FFN-2, :659-672 co-attention projections, :745-748 biOutput dense,
// :983 LM transform, :1025 tied decoder).  A is the activation matrix (bf16, K contiguous), W is the
// nn.Linear weight exactly as stored in the checkpoint ([out,in] = [N,K], K contiguous) cast to bf16,
// so both operands are "K-major" and no transpose is ever materialised.
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor 128x64 (A) and BNx64 (W) bf16 boxes, 128B swizzle,
//            into a kStages-deep shared-memory ring, completion on mbarriers
//   warp 1   one elected thread issues tcgen05.mma (M=128, N=BN, K=16) with the fp32 accumulator in
//            TMEM; tcgen05.commit releases ring slots and publishes finished accumulators
//   warp 2   TMEM allocator (2 x BN columns: the accumulator is double buffered so the epilogue of
//            tile i overlaps the main loop of tile i+1)
//   warps 4-7 epilogue: tcgen05.ld (each thread owns one accumulator row), + bias, GELU/ReLU,
//            + fp32 residual, stores fp32 and/or bf16;  or, in LSE mode (the fused LM head), an
//            online log-sum-exp over the tile's vocabulary columns and the label-logit pick, so the
//            [rows, 30522] logits never reach HBM.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace unimm {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;      // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;  // fixed for 16-bit inputs

template <int BN, int ST = 0, bool SM2 = false>
struct GemmCfg {
    static constexpr int kStages = ST > 0 ? ST : ((BN == 256 && !SM2) ? 4 : 6);
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (SM2 ? BN / 2 : BN) * BK * 2;     // a CTA pair keeps half of the W box in each CTA
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
};

// CL = 2: two CTAs of a cluster work on vertically adjacent 128-row tiles of the same BN columns and each loads only half
// of the shared W box, multicast into both CTAs' shared memory (operand traffic from L2 per CTA: 32 KB instead of 48 KB per
// k-block at BN = 256); a ring slot is reusable once BOTH CTAs' MMAs have read it (empty barriers count 2, multicast commit).
// CL = 3: the same two CTAs as ONE cta_group::2 MMA (M = 256 across the two SMs): each CTA holds only its half of the W box
// (no duplicate in shared memory, half the operand reads per SM), the leader CTA issues tcgen05.mma.cta_group::2 for both,
// both CTAs' TMA loads complete on the leader's full barrier, commits are multicast to both CTAs.
template <int BN, bool LSE, int ST = 0, bool FRAG = false, int CL = 1>
__global__ void __launch_bounds__(384, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 GemmEpilogue ep) {
    constexpr bool SM2 = CL == 3;
    constexpr int CS = CL == 1 ? 1 : 2;                      // CTAs per cluster
    using Cfg = GemmCfg<BN, ST, SM2>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need 1024-byte aligned bases
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::kStages * Cfg::kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float4* stage_all = reinterpret_cast<float4*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);   // 8 warps x 32x32 fp32

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = ((M + BM - 1) / BM + CS - 1) / CS;     // row blocks of CS x 128 rows (a padding tile is all out of bounds)
    const int num_n = (N + BN - 1) / BN;
    const int splits = ep.split_k > 1 ? ep.split_k : 1;      // split-K: consecutive tile indices share an output tile
    const int num_tiles = num_m * num_n * splits;
    const int num_k1 = (K + BK - 1) / BK;                    // K % 64 != 0 only with both operands MN-major (host-checked): zero-filled rows
    // fp32-class mode: every k-block three times — (a_lo, w_hi), (a_hi, w_lo), (a_hi, w_hi), small terms first — into one accumulator
    const int num_k = ep.split3 ? 3 * num_k1 : num_k1;
    const int k_per = (num_k + splits - 1) / splits;         // k-blocks per split (the host guarantees no split is empty)
    const uint32_t crank = CS > 1 ? ptx::cluster_ctarank() : 0u;
    const int first_tile = blockIdx.x / CS, tile_step = gridDim.x / CS;
    constexpr uint16_t kMask = static_cast<uint16_t>((1u << CS) - 1u);

    if (warp == 0 && ptx::elect_one()) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && ptx::elect_one()) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], CL == 2 ? 2 : 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], SM2 ? 16 : 8);  // one arrival per epilogue warp (of both CTAs for a CTA pair)
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if (SM2) ptx::tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
        else ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    ptx::tc_fence_before();
    if (CS > 1) ptx::cluster_sync_all();   // the peer's barriers exist before anything arrives on them remotely
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch_dependents();
    ptx::grid_dep_wait();                  // everything above overlapped the previous kernel's tail; A (and the residual) are its outputs

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                const int mn = tile / splits, sp = tile - mn * splits;
                const int m0 = ((mn / num_n) * CS + static_cast<int>(crank)) * BM;
                const int n0 = (mn % num_n) * BN;
                const int kb_end = min(num_k, (sp + 1) * k_per);
                for (int kb = sp * k_per; kb < kb_end; ++kb) {
                    int ka = kb * BK, kw = kb * BK;          // column of the A box / of the W box
                    if (ep.split3) {
                        const int pass = kb / num_k1, kk = (kb - pass * num_k1) * BK;
                        ka = kk + (pass == 0 ? K : 0);       // pass 0 reads the lo plane of A, pass 1 the lo plane of W
                        kw = kk + (pass == 1 ? K : 0);
                    }
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (SM2) {
                        // both CTAs' boxes are counted on the leader's barrier, which the leader arms for the pair
                        if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                        if (ep.a_mn) {       // MN-major: 64 x 64 boxes [contraction rows x 64 contiguous M / N elements], 8 KB atoms
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                ptx::tma_load_2d_2sm(sA + stage * Cfg::kABytes + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, ka);
                        } else {
                            ptx::tma_load_2d_2sm(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], ka, m0);
                        }
                        if (ep.b_mn) {
#pragma unroll
                            for (int j = 0; j < BN / 128; ++j)
                                ptx::tma_load_2d_2sm(sB + stage * Cfg::kBBytes + j * 8192, &tmB, &full_bar[stage],
                                                     n0 + static_cast<int>(crank) * (BN / 2) + 64 * j, kw);
                        } else {
                            ptx::tma_load_2d_2sm(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kw, n0 + static_cast<int>(crank) * (BN / 2));
                        }
                        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    if (ep.a_mn) {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            ptx::tma_load_2d(sA + stage * Cfg::kABytes + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, ka);
                    } else {
                        ptx::tma_load_2d(sA + stage * Cfg::kABytes, &tmA, &full_bar[stage], ka, m0);
                    }
                    if (CL == 1 && ep.b_mn) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            ptx::tma_load_2d(sB + stage * Cfg::kBBytes + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, kw);
                    } else if (CL == 1) ptx::tma_load_2d(sB + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kw, n0);
                    else ptx::tma_load_2d_mc(sB + stage * Cfg::kBBytes + crank * (Cfg::kBBytes / CL), &tmB, &full_bar[stage], kw,
                                             n0 + static_cast<int>(crank) * (BN / CL), kMask);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (ptx::elect_one() && (!SM2 || crank == 0)) {
            // bits 15 / 16 of the instruction descriptor: A / B operand is MN-major
            const uint32_t idesc = ptx::make_idesc_f16(SM2 ? 2 * BM : BM, BN, (ep.in_kind < 0 ? ep.lp_kind : ep.in_kind) == LP_FP16 ? 0u : 1u) | (ep.a_mn ? (1u << 15) : 0u) |
                                   (ep.b_mn ? (1u << 16) : 0u);
            // K-major: 16 elements along K = 32 bytes inside the swizzle atom (+2 in the addr >> 4 field); MN-major: 16 contraction
            // rows of 128 bytes = two 1024-byte groups (+128), 64-element atoms along M / N 8 KB apart (the leading byte offset)
            const uint32_t ka_step = ep.a_mn ? 128u : 2u, kb_step = ep.b_mn ? 128u : 2u;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                const int sp = tile % splits;
                const int kb_begin = sp * k_per, kb_end = min(num_k, (sp + 1) * k_per);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sA + stage * Cfg::kABytes), b_addr = ptx::smem_u32(sB + stage * Cfg::kBBytes);
                    const uint64_t da = ep.a_mn ? ptx::make_sw128_mnmajor_desc(a_addr, 8192) : ptx::make_sw128_kmajor_desc(a_addr);
                    const uint64_t db = ep.b_mn ? ptx::make_sw128_mnmajor_desc(b_addr, 8192) : ptx::make_sw128_kmajor_desc(b_addr);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t accum = (kb != kb_begin || k != 0) ? 1u : 0u;
                        if (SM2) ptx::umma_f16_ss_2sm(tmem_d, da + ka_step * k, db + kb_step * k, idesc, accum);
                        else ptx::umma_f16_ss(tmem_d, da + ka_step * k, db + kb_step * k, idesc, accum);
                    }
                    if (CL == 1) ptx::umma_commit(&empty_bar[stage]);  // slot reusable once these MMAs have read it
                    else if (CL == 2) ptx::umma_commit_mc(&empty_bar[stage], kMask);  // ... in both CTAs (the peer multicasts into this slot too)
                    else ptx::umma_commit_2sm_mc(&empty_bar[stage], kMask);   // both CTAs' producers may refill their halves
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (SM2) ptx::umma_commit_2sm_mc(&tfull_bar[acc], kMask);   // both CTAs' epilogues own half of the 256 rows
                else ptx::umma_commit(&tfull_bar[acc]);  // accumulator complete
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        // TMEM lane quarter = warp % 4; the two warps of a quarter split the tile's columns in halves.  With one
        // warp per scheduler the TMEM-load / shared-memory / global-load latencies of a chunk are exposed, two per
        // scheduler plus the one-chunk-ahead TMEM load below hide them.
        const int ew = (warp - 4) & 3;
        const int half = (warp - 4) >> 2;
        constexpr int HALF_N = BN / 2;
        constexpr int NCH = HALF_N / 32;
        float4* stage = stage_all + (warp - 4) * 256;
        // 128-bit global accesses need 16-byte aligned rows
        const bool fast_ok = (ep.out_f32 == nullptr || ((ep.ldo_f32 & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0)) &&
                             (ep.out_bf16 == nullptr || ((ep.ldo_bf16 & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.out_bf16) & 7) == 0)) &&
                             (ep.residual == nullptr || ((ep.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0)) &&
                             (ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0);
        // 16-bit-only output with 128-byte aligned row segments: the paired-chunk path (needs N % 64 == 0 so pairs never split)
        const bool lp_only = fast_ok && !ep.out_hilo && ep.out_bf16 != nullptr && ep.out_f32 == nullptr && ep.residual == nullptr && (N % 64) == 0 &&
                             (ep.ldo_bf16 & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) == 0;
        // FRAG instantiation: the host has checked lp_only && w_perm16 && N % 32 == 0 (launch()); the other paths are compiled out
        // the accumulator pair is handed back to the MMA warp of the LEADER CTA
        auto arrive_tempty = [&](uint64_t* bar) {
            if (SM2 && crank != 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
            else ptx::mbar_arrive(bar);
        };
        int acc = 0;
        uint32_t acc_phase = 0;
        float out_amax = 0.f;                  // ep.amax_out: running max |out| of this thread's stores
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            const int mn = tile / splits;
            const int tn = mn % num_n;
            const int m0 = ((mn / num_n) * CS + static_cast<int>(crank)) * BM;
            const int n0 = tn * BN + half * HALF_N;
            const int row = m0 + ew * 32 + lane;
            const bool row_ok = row < M;
            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr0 = tmem_base + acc * BN + half * HALF_N + (static_cast<uint32_t>(ew * 32) << 16);

            float run_max = -INFINITY, run_sum = 0.f, lab_logit = 0.f;
            int label = -1;
            if (LSE && row_ok) label = ep.labels[row];

            if constexpr (FRAG) {
                // ---- 16-bit-only outputs with fragment-ordered weights (QKV, FFN-1): tcgen05.ld.16x256b hands each thread
                // 8 CONSECUTIVE output columns of 4 rows (see permute_weight_rows, mode 1), so bias + activation + pack + one
                // 16-byte store per row happen straight out of the registers: no shared-memory transpose, no __syncwarp
                const int a4 = lane & 3, r8 = lane >> 2;
                uint32_t f[2][32];
                auto ld_chunk = [&](int c, uint32_t* dst) {
                    ptx::tmem_ld_16x256b_x4(taddr0 + c * 32, dst);
                    ptx::tmem_ld_16x256b_x4(taddr0 + c * 32 + (16u << 16), dst + 16);
                };
                if (n0 < N) ld_chunk(0, f[0]);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int col0 = n0 + c * 32;
                    if (col0 >= N) break;  // warp-uniform (N % 32 == 0)
                    float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (ep.bias != nullptr) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * a4));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + 8 * a4 + 4));
                        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
                    }
                    ptx::tmem_ld_wait();
                    if (c + 1 < NCH && col0 + 32 < N) ld_chunk(c + 1, f[(c + 1) & 1]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {          // row 8 k + r8 of the warp's 32: slab k >> 1, half k & 1
                        float y[8];
#pragma unroll
                        for (int m = 0; m < 8; m += 2) {
                            const int reg = (k >> 1) * 16 + 4 * (m >> 1) + 2 * (k & 1);
                            f32x2::unpack(f32x2::add(f32x2::pack(__uint_as_float(f[c & 1][reg]), __uint_as_float(f[c & 1][reg + 1])),
                                                     f32x2::pack(bb[m], bb[m + 1])), y[m], y[m + 1]);
                        }
                        const int grow = m0 + ew * 32 + 8 * k + r8;
                        if (ep.pre_act_lp && grow < M) {    // training forward: the activation's input as 16-bit values (out_f32 points at them)
                            uint4 pp;
                            pp.x = pack_lp2(y[0], y[1], ep.lp_kind); pp.y = pack_lp2(y[2], y[3], ep.lp_kind);
                            pp.z = pack_lp2(y[4], y[5], ep.lp_kind); pp.w = pack_lp2(y[6], y[7], ep.lp_kind);
                            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out_f32) + static_cast<size_t>(grow) * ep.ldo_f32 + col0 + 8 * a4) = pp;
                        }
                        if (ep.act == ACT_GELU) {
#pragma unroll
                            for (int m = 0; m < 8; m += 2) gelu_fast2(y[m], y[m + 1]);
                        } else if (ep.act == ACT_GELU_TANH) {
#pragma unroll
                            for (int m = 0; m < 8; m += 2) gelu_tanh2(y[m], y[m + 1]);
                        } else if (ep.act == ACT_RELU) {
#pragma unroll
                            for (int m = 0; m < 8; ++m) y[m] = fmaxf(y[m], 0.f);
                        }
                        uint4 pk;
                        pk.x = pack_lp2(y[0], y[1], ep.lp_kind); pk.y = pack_lp2(y[2], y[3], ep.lp_kind);
                        pk.z = pack_lp2(y[4], y[5], ep.lp_kind); pk.w = pack_lp2(y[6], y[7], ep.lp_kind);
                        if (grow < M) *reinterpret_cast<uint4*>(ep.out_bf16 + static_cast<size_t>(grow) * ep.ldo_bf16 + col0 + 8 * a4) = pk;
                    }
                }
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_tempty(&tempty_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
                continue;
            }
            if constexpr (FRAG) continue;   // (unreachable; keeps the legacy paths out of this instantiation)
            uint32_t v[32];
            if (ep.debug_mode == 3) goto tile_done;
            if (n0 < N) ptx::tmem_ld_32x32b_x32(taddr0, v);      // chunk 0 in flight
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= N) break;  // warp-uniform
                const int ncols = min(32, N - col0);
                ptx::tmem_ld_wait();
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
                if (c + 1 < NCH && col0 + 32 < N) ptx::tmem_ld_32x32b_x32(taddr0 + (c + 1) * 32, v);   // next chunk in flight
                if (LSE) {
                    constexpr float kLog2e = 1.4426950408889634f;
                    if (ep.dz != nullptr) {
                        // ---- LM-head backward: dz = coef * (softmax - onehot) recomputed from the saved log-sum-exp, written in both
                        // orientations as 16-bit GEMM operands (row-major for dH = dz E, transposed for dE = dz^T h)
                        const float l2 = row_ok ? ep.lse[row] * kLog2e : 0.f;
                        const float cf = row_ok ? ep.coef[row] : 0.f;
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float d0 = 0.f, d1 = 0.f;
                            if (j < ncols) {
                                const float v = x[j] + __ldg(ep.bias + col0 + j);
                                d0 = cf * (fast_ex2(fmaf(v, kLog2e, -l2)) - (col0 + j == label ? 1.f : 0.f));
                            }
                            if (j + 1 < ncols) {
                                const float v = x[j + 1] + __ldg(ep.bias + col0 + j + 1);
                                d1 = cf * (fast_ex2(fmaf(v, kLog2e, -l2)) - (col0 + j + 1 == label ? 1.f : 0.f));
                            }
                            pk[j >> 1] = pack_lp2(d0, d1, ep.lp_kind);
                        }
                        if (row_ok) {
                            bf16* drow = ep.dz + static_cast<size_t>(row) * ep.ldz + col0;
                            if (col0 + 32 <= ep.dz_cols) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    *reinterpret_cast<uint4*>(drow + 8 * q) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; j += 2)
                                    if (col0 + j + 2 <= ep.dz_cols) *reinterpret_cast<uint32_t*>(drow + j) = pk[j >> 1];
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) {
                                    const uint16_t h16 = static_cast<uint16_t>((j & 1) ? (pk[j >> 1] >> 16) : (pk[j >> 1] & 0xffffu));
                                    reinterpret_cast<uint16_t*>(ep.dzT)[static_cast<size_t>(col0 + j) * ep.ldzt + row] = h16;
                                }
                        }
                        continue;
                    }
                    if (ncols == 32 && fast_ok) {
                        // full chunk (all but the last vocabulary tile): 128-bit uniform bias loads, no per-column predicates
                        if (ep.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                                x[j] += b4.x; x[j + 1] += b4.y; x[j + 2] += b4.z; x[j + 3] += b4.w;
                            }
                        }
                        float c0 = fmaxf(x[0], x[1]), c1 = fmaxf(x[2], x[3]);
#pragma unroll
                        for (int j = 4; j < 32; j += 4) { c0 = fmaxf(c0, fmaxf(x[j], x[j + 1])); c1 = fmaxf(c1, fmaxf(x[j + 2], x[j + 3])); }
                        const float nmax = fmaxf(run_max, fmaxf(c0, c1));
                        const float nm2 = nmax * kLog2e;
                        float s0 = run_sum * fast_ex2((run_max - nmax) * kLog2e), s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            s0 += fast_ex2(fmaf(x[j], kLog2e, -nm2));
                            s1 += fast_ex2(fmaf(x[j + 1], kLog2e, -nm2));
                            s2 += fast_ex2(fmaf(x[j + 2], kLog2e, -nm2));
                            s3 += fast_ex2(fmaf(x[j + 3], kLog2e, -nm2));
                        }
                        const int lj = label - col0;
                        if (lj >= 0 && lj < 32) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j == lj) lab_logit = x[j];
                        }
                        run_max = nmax;
                        run_sum = (s0 + s1) + (s2 + s3);
                        continue;
                    }
                    if (ep.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) x[j] += __ldg(ep.bias + col0 + j);
                    }
                    float cmax = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) cmax = fmaxf(cmax, x[j]);
                    const float nmax = fmaxf(run_max, cmax);
                    const float nm2 = nmax * kLog2e;
                    float s = run_sum * fast_ex2((run_max - nmax) * kLog2e);  // 2^-inf = 0 on the first chunk
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) {
                            s += fast_ex2(fmaf(x[j], kLog2e, -nm2));           // one FFMA + one SFU op per logit
                            if (col0 + j == label) lab_logit = x[j];
                        }
                    run_max = nmax;
                    run_sum = s;
                    continue;
                }
                if (ep.debug_mode == 2) { if (x[0] == 1.2345e30f) ep.out_f32[0] = x[5]; continue; }
                if (lp_only && ncols == 32 && ep.debug_mode != 1) {
                    // ---- 16-bit-only outputs (QKV, FFN-1): bias + activation in the row-per-thread layout, pack to 16 bit,
                    // stage TWO chunks (64 columns = 128 bytes per row) so every global store writes whole 128-byte lines
                    if (ep.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                            f32x2::unpack(f32x2::add(f32x2::pack(x[j], x[j + 1]), f32x2::pack(b4.x, b4.y)), x[j], x[j + 1]);
                            f32x2::unpack(f32x2::add(f32x2::pack(x[j + 2], x[j + 3]), f32x2::pack(b4.z, b4.w)), x[j + 2], x[j + 3]);
                        }
                    }
                    // activation hoisted out of the element loop: a branch-free body lets the 32 independent chains interleave
                    if (ep.act == ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) gelu_epi2(x[j], x[j + 1]);
                    } else if (ep.act == ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                    }
                    uint4* stage16 = reinterpret_cast<uint4*>(stage);
                    const int slot0 = (c & 1) * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 pk;
                        pk.x = pack_lp2(x[8 * q], x[8 * q + 1], ep.lp_kind); pk.y = pack_lp2(x[8 * q + 2], x[8 * q + 3], ep.lp_kind);
                        pk.z = pack_lp2(x[8 * q + 4], x[8 * q + 5], ep.lp_kind); pk.w = pack_lp2(x[8 * q + 6], x[8 * q + 7], ep.lp_kind);
                        stage16[lane * 8 + ((slot0 + q) ^ (lane & 7))] = pk;
                    }
                    if (c & 1) {
                        __syncwarp();
                        const int sl = lane & 7;
                        const int cc = col0 - 32 + sl * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = (lane >> 3) + 4 * i;
                            const int grow = m0 + ew * 32 + r;
                            const uint4 pk = stage16[r * 8 + (sl ^ (r & 7))];
                            if (grow < M) *reinterpret_cast<uint4*>(ep.out_bf16 + static_cast<size_t>(grow) * ep.ldo_bf16 + cc) = pk;
                        }
                        __syncwarp();
                    }
                    continue;
                }
                if (fast_ok && ncols == 32 && ep.debug_mode != 1) {
                    // ---- coalesced path: transpose the warp's 32x32 fp32 block through swizzled shared memory so that
                    // 8 lanes cover 128 contiguous bytes of one output row (4 rows per warp instruction)
                    const int ch = lane & 7;
                    const int cc = col0 + ch * 4;
                    const int r0 = m0 + ew * 32 + (lane >> 3);
                    // residual / bias loads do not depend on the accumulator: issue all of them first (8 independent
                    // 128-bit loads in flight per thread) so their latency overlaps the shared-memory transpose
                    float4 res[8];
                    if (ep.residual != nullptr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int grow = r0 + 4 * i;
                            res[i] = (grow < M) ? *reinterpret_cast<const float4*>(ep.residual + static_cast<size_t>(grow) * ep.ldr + cc)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + cc));
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8)
                        stage[lane * 8 + (c8 ^ (lane & 7))] = make_float4(x[4 * c8], x[4 * c8 + 1], x[4 * c8 + 2], x[4 * c8 + 3]);
                    __syncwarp();
                    float4 y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = (lane >> 3) + 4 * i;
                        y[i] = stage[r * 8 + (ch ^ (r & 7))];
                    }
                    __syncwarp();   // the staging block is rewritten by the next chunk
                    const float alpha = ep.alpha_ptr != nullptr ? __ldg(ep.alpha_ptr) : ep.alpha;
                    if (alpha != 1.f) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { y[i].x *= alpha; y[i].y *= alpha; y[i].z *= alpha; y[i].w *= alpha; }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { y[i].x += b4.x; y[i].y += b4.y; y[i].z += b4.z; y[i].w += b4.w; }
                    if (ep.drop.thresh != 0u) {             // dropout on the projection's output, before the residual joins
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t i0 = static_cast<uint32_t>(r0 + 4 * i) * static_cast<uint32_t>(N) + static_cast<uint32_t>(cc);
                            y[i].x = drop_keep(ep.drop.seed, i0, ep.drop.thresh) ? y[i].x * ep.drop.scale : 0.f;
                            y[i].y = drop_keep(ep.drop.seed, i0 + 1u, ep.drop.thresh) ? y[i].y * ep.drop.scale : 0.f;
                            y[i].z = drop_keep(ep.drop.seed, i0 + 2u, ep.drop.thresh) ? y[i].z * ep.drop.scale : 0.f;
                            y[i].w = drop_keep(ep.drop.seed, i0 + 3u, ep.drop.thresh) ? y[i].w * ep.drop.scale : 0.f;
                        }
                    }
                    if (ep.pre_act_f32 && ep.pre_act_lp) {      // the pre-activation as 16-bit values (out_f32 points at them)
                        bf16* pre = reinterpret_cast<bf16*>(ep.out_f32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int grow = r0 + 4 * i;
                            if (grow < M)
                                *reinterpret_cast<uint2*>(pre + static_cast<size_t>(grow) * ep.ldo_f32 + cc) =
                                    make_uint2(pack_lp2(y[i].x, y[i].y, ep.lp_kind), pack_lp2(y[i].z, y[i].w, ep.lp_kind));
                        }
                    } else if (ep.pre_act_f32) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int grow = r0 + 4 * i;
                            if (grow < M) *reinterpret_cast<float4*>(ep.out_f32 + static_cast<size_t>(grow) * ep.ldo_f32 + cc) = y[i];
                        }
                    }
                    if (ep.act == ACT_GELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            gelu_epi2(y[i].x, y[i].y); gelu_epi2(y[i].z, y[i].w);
                        }
                    } else if (ep.act == ACT_GELU_ERF) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            y[i].x = gelu_erf(y[i].x); y[i].y = gelu_erf(y[i].y); y[i].z = gelu_erf(y[i].z); y[i].w = gelu_erf(y[i].w);
                        }
                    } else if (ep.act == ACT_RELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            y[i].x = fmaxf(y[i].x, 0.f); y[i].y = fmaxf(y[i].y, 0.f); y[i].z = fmaxf(y[i].z, 0.f); y[i].w = fmaxf(y[i].w, 0.f);
                        }
                    }
                    if (ep.residual != nullptr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { y[i].x += res[i].x; y[i].y += res[i].y; y[i].z += res[i].z; y[i].w += res[i].w; }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int grow = r0 + 4 * i;
                        if (grow < M) {
                            if (ep.amax_out != nullptr)
                                out_amax = fmaxf(fmaxf(out_amax, fmaxf(fabsf(y[i].x), fabsf(y[i].y))), fmaxf(fabsf(y[i].z), fabsf(y[i].w)));
                            if (ep.out_f32 != nullptr && splits > 1) {
                                float* o = ep.out_f32 + static_cast<size_t>(grow) * ep.ldo_f32 + cc;
                                atomicAdd(o, y[i].x); atomicAdd(o + 1, y[i].y); atomicAdd(o + 2, y[i].z); atomicAdd(o + 3, y[i].w);
                            } else if (ep.out_f32 != nullptr && !ep.pre_act_f32)
                                *reinterpret_cast<float4*>(ep.out_f32 + static_cast<size_t>(grow) * ep.ldo_f32 + cc) = y[i];
                            if (ep.out_bf16 != nullptr && ep.out_hilo) {
                                uint2 hi, lo;
                                split_hilo2(y[i].x, y[i].y, hi.x, lo.x);
                                split_hilo2(y[i].z, y[i].w, hi.y, lo.y);
                                *reinterpret_cast<uint2*>(ep.out_bf16 + static_cast<size_t>(grow) * ep.ldo_bf16 + cc) = hi;
                                *reinterpret_cast<uint2*>(ep.out_bf16 + static_cast<size_t>(grow) * ep.ldo_bf16 + (ep.hilo_off > 0 ? ep.hilo_off : N) + cc) = lo;
                            } else if (ep.out_bf16 != nullptr) {
                                uint2 pk;
                                pk.x = pack_lp2(y[i].x, y[i].y, ep.lp_kind);
                                pk.y = pack_lp2(y[i].z, y[i].w, ep.lp_kind);
                                *reinterpret_cast<uint2*>(ep.out_bf16 + static_cast<size_t>(grow) * ep.ldo_bf16 + cc) = pk;
                            }
                        }
                    }
                    continue;
                }
                // ---- generic path (ragged last columns, unaligned leading dimensions): one row per thread
                {
                    const float alpha = ep.alpha_ptr != nullptr ? __ldg(ep.alpha_ptr) : ep.alpha;
                    if (alpha != 1.f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] *= alpha;
                    }
                }
                if (ep.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) x[j] += __ldg(ep.bias + col0 + j);
                }
                if (ep.drop.thresh != 0u) {
                    const uint32_t i0 = static_cast<uint32_t>(row) * static_cast<uint32_t>(N) + static_cast<uint32_t>(col0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = drop_keep(ep.drop.seed, i0 + j, ep.drop.thresh) ? x[j] * ep.drop.scale : 0.f;
                }
                if (ep.pre_act_f32 && ep.pre_act_lp && row_ok) {
                    bf16* o = reinterpret_cast<bf16*>(ep.out_f32) + static_cast<size_t>(row) * ep.ldo_f32 + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) o[j] = lp_from_f32(x[j], ep.lp_kind);
                } else if (ep.pre_act_f32 && row_ok) {
                    float* o = ep.out_f32 + static_cast<size_t>(row) * ep.ldo_f32 + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) o[j] = x[j];
                }
                if (ep.act == ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_fast(x[j]);
                } else if (ep.act == ACT_GELU_ERF) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
                } else if (ep.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (!row_ok) continue;
                if (ep.residual != nullptr) {
                    const float* r = ep.residual + static_cast<size_t>(row) * ep.ldr + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) x[j] += r[j];
                }
                if (ep.amax_out != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) out_amax = fmaxf(out_amax, fabsf(x[j]));
                }
                if (ep.out_f32 != nullptr && !ep.pre_act_f32) {
                    float* o = ep.out_f32 + static_cast<size_t>(row) * ep.ldo_f32 + col0;
                    if (splits > 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) atomicAdd(o + j, x[j]);
                    } else if (fast_ok && ncols == 32) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) o[j] = x[j];
                    }
                }
                if (ep.out_bf16 != nullptr) {
                    bf16* o = ep.out_bf16 + static_cast<size_t>(row) * ep.ldo_bf16 + col0;
                    if (fast_ok && ncols == 32 && (ep.ldo_bf16 & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 pk;
                            pk.x = pack_lp2(x[j], x[j + 1], ep.lp_kind); pk.y = pack_lp2(x[j + 2], x[j + 3], ep.lp_kind);
                            pk.z = pack_lp2(x[j + 4], x[j + 5], ep.lp_kind); pk.w = pack_lp2(x[j + 6], x[j + 7], ep.lp_kind);
                            *reinterpret_cast<uint4*>(o + j) = pk;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) o[j] = lp_from_f32(x[j], ep.lp_kind);
                    }
                }
            }
        tile_done:
            ptx::tmem_ld_wait();   // no TMEM load may be outstanding when the accumulator is handed back
            if (LSE && ep.dz != nullptr) {
                // nothing per tile: dz went out chunk by chunk
            } else if (LSE && row_ok && n0 < N) {
                ep.partials[static_cast<size_t>(row) * (2 * num_n) + 2 * tn + half] = make_float2(run_max, run_sum);
                if (label >= n0 && label < min(n0 + HALF_N, N)) ep.label_logit[row] = lab_logit;
            } else if (LSE && row_ok) {
                ep.partials[static_cast<size_t>(row) * (2 * num_n) + 2 * tn + half] = make_float2(-INFINITY, 0.f);
            }
            // hand the accumulator back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (ep.amax_out != nullptr) {
            const unsigned u = __reduce_max_sync(0xffffffffu, __float_as_uint(out_amax));
            if (lane == 0 && u != 0u) atomicMax(ep.amax_out, u);
        }
    }
    ptx::tc_fence_before();
    if (CS > 1) ptx::cluster_sync_all();   // no CTA leaves while its peer may still multicast into its shared memory / barriers
    else __syncthreads();
    if (warp == 2) {
        if (SM2) ptx::tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
        else ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched through the runtime, so the library
// does not link libcuda) and a small cache keyed by (pointer, shape, stride, box).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    int rows, cols, ld, box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (size_t)k.cols * 0xC2B2AE3D27D4EB4Full + (size_t)k.ld * 0x165667B19E3779F9ull +
             (size_t)k.box_rows;
        return h;
    }
};

int make_map_bf16(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = get_encode_fn();
    UNIMM_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    UNIMM_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0, "TMA operand must be 16-byte aligned");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    // bf16 and fp16 tiles are both plain 2-byte elements to TMA (no arithmetic, zero OOB fill)
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNIMM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
    return 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int BN, bool LSE, int ST = 0, bool FRAG = false, int CL = 1>
int launch(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep, int max_ctas,
           cudaStream_t stream) {
    constexpr int CS = CL == 1 ? 1 : 2;
    using Cfg = GemmCfg<BN, ST, CL == 3>;
    CUtensorMap tmA, tmB;
    const int kcols = ep.split3 ? 2 * K : K;                 // fp32-class mode: hi | lo planes side by side
    // MN-major operands are [K, M] / [K, N] matrices read in 64 x 64 boxes
    if (ep.a_mn) UNIMM_TRY(make_map_bf16(A, K, M, lda, 64, &tmA));
    else UNIMM_TRY(make_map_bf16(A, M, kcols, lda, BM, &tmA));
    if (ep.b_mn) UNIMM_TRY(make_map_bf16(W, K, N, ldw, 64, &tmB));
    else UNIMM_TRY(make_map_bf16(W, N, kcols, ldw, BN / CS, &tmB));
    static int max_clusters = 0;
    auto kernel = umma_gemm_kernel<BN, LSE, ST, FRAG, CL>;
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), Cfg::kSmemBytes));
    const int tiles = (((M + BM - 1) / BM + CS - 1) / CS) * ((N + BN - 1) / BN) * (ep.split_k > 1 ? ep.split_k : 1);
    if (CL == 1) {
        int grid = tiles < num_sms() ? tiles : num_sms();
        if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
        cudaLaunchAttribute attr1[1];
        unsigned n1 = 0;
        add_pdl_attr(attr1, &n1);
        cudaLaunchConfig_t cfg1 = {};
        cfg1.gridDim = dim3(grid, 1, 1);
        cfg1.blockDim = dim3(384, 1, 1);
        cfg1.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg1.stream = stream;
        cfg1.attrs = attr1;
        cfg1.numAttrs = n1;
        UNIMM_CUDA_CHECK(cudaLaunchKernelEx(&cfg1, kernel, tmA, tmB, M, N, K, ep));
        UNIMM_LAUNCH_CHECK(1);
        return 0;
    }
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters == 0) {
        cfg.gridDim = dim3(num_sms() / CS * CS, 1, 1);
        int n = 0;
        UNIMM_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
        UNIMM_CHECK(n > 0, "no co-resident cluster fits the multicast GEMM");
        max_clusters = n < num_sms() / CS ? n : num_sms() / CS;
    }
    int clusters = tiles < max_clusters ? tiles : max_clusters;
    if (max_ctas > 0 && clusters > max_ctas / CS) clusters = max_ctas / CS > 0 ? max_ctas / CS : 1;
    cfg.gridDim = dim3(clusters * CS, 1, 1);
    add_pdl_attr(attr, &cfg.numAttrs);
    UNIMM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, M, N, K, ep));
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace

int gemm_make_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
    return make_map_bf16(ptr, rows, cols, ld, box_rows, out);
}
int gemm_num_sms() { return num_sms(); }

int gemm_umma_bf16(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep, int tile_n,
                   int max_ctas, cudaStream_t stream) {
    UNIMM_CHECK(M > 0 && N > 0 && K > 0 && (K % BK == 0 || (ep.a_mn && ep.b_mn)), "umma gemm: K must be a positive multiple of 64");
    const bool lse = ep.partials != nullptr || ep.dz != nullptr;
    UNIMM_CHECK(!ep.pre_act_f32 || (ep.out_f32 != nullptr && ep.out_bf16 != nullptr && ep.residual == nullptr && !lse && (!ep.w_perm16 || ep.pre_act_lp) &&
                                    !ep.out_hilo && ep.split_k <= 1 && ep.amax_out == nullptr),
                "pre-activation output: fp32 pre-activation + 16-bit activation, plain epilogue, no residual");
    UNIMM_CHECK(!ep.pre_act_lp || (ep.pre_act_f32 && ep.ldo_f32 % 4 == 0), "16-bit pre-activation: a variant of the pre-activation epilogue, rows 8-byte aligned");
    UNIMM_CHECK(ep.drop.thresh == 0u || (!lse && !ep.w_perm16 && !ep.out_hilo && ep.act == ACT_NONE && ep.split_k <= 1 && ep.out_f32 != nullptr &&
                                         ep.out_bf16 == nullptr && static_cast<double>(M) * N < 4294967296.0),
                "output dropout: plain fp32-output epilogue without activation");
    if (ep.a_mn || ep.b_mn || ep.split_k > 1) {
        UNIMM_CHECK(!lse && !ep.split3 && !ep.w_perm16 && ep.debug_mode == 0, "MN-major operands / split-K: plain epilogue only");
        UNIMM_CHECK(!ep.a_mn || ((lda & 7) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0), "MN-major A: 16-byte aligned rows");
        UNIMM_CHECK(!ep.b_mn || ((ldw & 7) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0), "MN-major W: 16-byte aligned rows");
        GemmEpilogue e2 = ep;
        if (ep.split_k > 1) {
            UNIMM_CHECK(ep.out_f32 != nullptr && ep.out_bf16 == nullptr && ep.bias == nullptr && ep.residual == nullptr && ep.act == ACT_NONE,
                        "split-K accumulates plain fp32 partial products");
            const int num_k = (K + BK - 1) / BK;
            int sk = ep.split_k < num_k ? ep.split_k : num_k;
            const int k_per = (num_k + sk - 1) / sk;
            e2.split_k = (num_k + k_per - 1) / k_per;          // no empty split
        }
        if (tile_n == 0) tile_n = (N % 256 == 0 || N > 2048) ? 256 : 128;
        const bool pair = M >= 8192 && tile_n == 256;          // tall problems: cta_group::2 pairs (CL = 2's multicast path is K-major only)
        if (pair) return launch<256, false, 0, false, 3>(A, lda, W, ldw, M, N, K, e2, max_ctas, stream);
        if (tile_n == 256) return launch<256, false>(A, lda, W, ldw, M, N, K, e2, max_ctas, stream);
        return launch<128, false>(A, lda, W, ldw, M, N, K, e2, max_ctas, stream);
    }
    UNIMM_CHECK(ep.dz == nullptr || (ep.dzT != nullptr && ep.lse != nullptr && ep.coef != nullptr && ep.labels != nullptr && ep.bias != nullptr &&
                                     (ep.ldz & 7) == 0 && (ep.dz_cols & 1) == 0 && ep.dz_cols <= ep.ldz && ep.ldzt >= M),
                "dz epilogue: incomplete arguments");
    UNIMM_CHECK((ep.alpha == 1.f && ep.alpha_ptr == nullptr) || (!lse && ep.out_bf16 == nullptr), "alpha applies to the fp32-output epilogue only");
    UNIMM_CHECK(!ep.split3 || (!ep.w_perm16 && ep.lp_kind == LP_FP16), "split3 takes fp16 hi | lo planes (plain or log-sum-exp epilogue)");
    UNIMM_CHECK(!ep.out_hilo || (ep.out_bf16 != nullptr && N % 32 == 0 && (ep.ldo_bf16 & 3) == 0 && (ep.hilo_off & 3) == 0 &&
                                 ep.ldo_bf16 >= (ep.hilo_off > 0 ? ep.hilo_off : N) + N && (reinterpret_cast<uintptr_t>(ep.out_bf16) & 7) == 0),
                "hi | lo output needs N % 32 == 0 and 8-byte aligned planes inside the row");
    if (tile_n == 0) tile_n = (N % 256 == 0 || N > 2048) ? 256 : 128;
    // tall problems (every SM gets several row blocks): CTA pairs sharing the W tile by TMA multicast
    // UNIMM_GEMM_MULTICAST: 0 = independent CTAs, 1 = CTA pairs sharing W by multicast, 2 = CTA pairs as one cta_group::2 MMA
    // (measured, profiles/r01_v8: mode 2 is +3..9 % per projection GEMM over mode 1 and on par with / above cuBLAS; the LM head's
    // log-sum-exp epilogue is the slower side there and prefers mode 1)
    static const int pair_mode = getenv("UNIMM_GEMM_MULTICAST") == nullptr ? 2 : atoi(getenv("UNIMM_GEMM_MULTICAST"));
    const bool tall = M >= 8192 && ep.debug_mode == 0;
    const bool mc = pair_mode == 1 && tall;
    const bool sm2 = pair_mode == 2 && tall;
    if (lse) {
        UNIMM_CHECK(tile_n == 256, "LSE epilogue uses 256-wide vocabulary tiles");
        static const bool lse_sm2 = getenv("UNIMM_LSE_SM2") != nullptr && atoi(getenv("UNIMM_LSE_SM2")) != 0;
        if (sm2 && lse_sm2) return launch<256, true, 0, false, 3>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
        if (mc || sm2) return launch<256, true, 0, false, 2>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
        return launch<256, true>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
    }
    if (tile_n == 256 && ep.debug_mode >= 4) {   // microbenchmark: 3-stage ring, epilogue mode = debug_mode - 4
        GemmEpilogue e2 = ep;
        e2.debug_mode -= 4;
        return launch<256, false, 3>(A, lda, W, ldw, M, N, K, e2, max_ctas, stream);
    }
    if (ep.w_perm16) {
        auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        UNIMM_CHECK(!ep.pre_act_lp || ((ep.ldo_f32 & 7) == 0 && a16(ep.out_f32)), "16-bit pre-activation of the fragment epilogue: 16-byte aligned rows");
        UNIMM_CHECK(ep.out_bf16 != nullptr && (ep.out_f32 == nullptr || ep.pre_act_lp) && ep.residual == nullptr && N % 32 == 0 && (ep.ldo_bf16 & 7) == 0 &&
                        a16(ep.out_bf16) && (ep.bias == nullptr || a16(ep.bias)) && ep.debug_mode == 0,
                    "fragment-ordered weights need a 16-bit-only, 16-byte aligned output with N % 32 == 0");
        if (tile_n == 256 && mc) return launch<256, false, 0, true, 2>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
        if (tile_n == 256 && sm2) return launch<256, false, 0, true, 3>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
        if (tile_n == 256) return launch<256, false, 0, true>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
        return launch<128, false, 0, true>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
    }
    if (tile_n == 256 && mc) return launch<256, false, 0, false, 2>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
    if (tile_n == 256 && sm2) return launch<256, false, 0, false, 3>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
    if (tile_n == 256) return launch<256, false>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
    UNIMM_CHECK(tile_n == 128, "umma gemm: tile_n must be 128 or 256");
    return launch<128, false>(A, lda, W, ldw, M, N, K, ep, max_ctas, stream);
}

int gemm_umma_lse_tiles(int N) { return 2 * ((N + 255) / 256); }   // two column halves per 256-wide tile

}  // namespace unimm
