// Fused masked attention on the tensor cores (bf16 in, fp32 softmax/accumulate, bf16 out).
//
// Same contract and mask generation as attention.cu (the CUDA-core kernel it is tested against); this
// is the throughput version used by the bf16 engine mode.  One CTA = 4 warps = 64 query rows of one
// (sequence, head); each warp owns 16 rows.  The keys/values the rows may attend (the union of their
// descriptor intervals, rounded to 64) are staged once in shared memory with cp.async, then every
// warp runs a FlashAttention-2 style loop over 64-key tiles: S = Q K^T with mma.sync.m16n8k16 (bf16),
// scale + on-the-fly mask + online softmax in registers, P re-used as the A fragment of O += P V.
// Attention is ~4 % of the path's FLOPs; the projections around it run on tcgen05.
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

constexpr int MQT = 64;   // query rows per CTA
constexpr int MKT = 64;   // keys per inner tile
constexpr int PADE = 8;   // bf16 elements of row padding: 16 B shifts successive rows by 4 banks (ldmatrix conflict-free)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr(p)));
}
template <bool FP16>
__device__ __forceinline__ void mma_lp(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (FP16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { return FP16 ? pack_fp16x2(lo, hi) : pack_bf16x2(lo, hi); }
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D, bool FP16>
__global__ void __launch_bounds__(128)
attn_mma_kernel(AttnArgs a, int kv_rows_max) {
    constexpr int LD = D + PADE;       // smem row stride in elements
    extern __shared__ __align__(16) uint8_t smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);          // [MQT][LD]
    bf16* Ks = Qs + MQT * LD;                              // [kv_rows_max][LD]
    bf16* Vs = Ks + static_cast<size_t>(kv_rows_max) * LD; // [kv_rows_max][LD]
    float* Ms = reinterpret_cast<float*>(Vs + static_cast<size_t>(kv_rows_max) * LD);  // [kv_rows_max] key mask (KEY_VECTOR)
    __shared__ int s_any_key;

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MQT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int Sq = a.Sq, Skv = a.Skv;
    const bf16* Q = static_cast<const bf16*>(a.q) + static_cast<size_t>(b) * Sq * a.ldq + h * D;
    const bf16* K = static_cast<const bf16*>(a.k) + static_cast<size_t>(b) * Skv * a.ldk + h * D;
    const bf16* V = static_cast<const bf16*>(a.v) + static_cast<size_t>(b) * Skv * a.ldv + h * D;
    bf16* O = static_cast<bf16*>(a.o) + static_cast<size_t>(b) * Sq * a.ldo + h * D;

    // ---- allowed key sets: the two rows this thread holds accumulators for, and the CTA-wide union
    SeqDesc desc = {0, 0, 0, 0};
    if (a.mask_kind != MASK_KEY_VECTOR) desc = a.desc[b];
    int lo[2] = {0, 0}, hi[2] = {Skv, Skv}, self[2] = {-1, -1};
    int kv_lo = 0, kv_hi = Skv;
    if (a.mask_kind == MASK_TEXT_SELF) {
        kv_lo = Skv; kv_hi = 0;
        for (int r = 0; r < MQT && q0 + r < Sq; ++r) {
            int l, hh, s;
            text_row_interval(desc, q0 + r, Skv, l, hh, s);
            if (hh <= l && s < 0) { l = 0; hh = Skv; }   // padding row: attends everything (see attention.cu)
            kv_lo = min(kv_lo, l);
            kv_hi = max(kv_hi, max(hh, s + 1));
            if (r == warp * 16 + g) { lo[0] = l; hi[0] = hh; self[0] = s; }
            if (r == warp * 16 + g + 8) { lo[1] = l; hi[1] = hh; self[1] = s; }
        }
    } else if (a.mask_kind == MASK_CO_INTERVAL) {
        int l, hh;
        co_interval(desc, Skv, l, hh);
        if (hh <= l) { l = 0; hh = Skv; }
        lo[0] = lo[1] = kv_lo = l;
        hi[0] = hi[1] = kv_hi = hh;
    }
    const int t_begin = (kv_lo / MKT) * MKT;
    const int t_end = ((kv_hi + MKT - 1) / MKT) * MKT;      // <= kv_rows_max + t_begin by construction
    const int n_rows = t_end - t_begin;

    // ---- stage Q tile and the K/V rows [t_begin, t_end) (rows >= Skv are zero filled)
    if (tid == 0) s_any_key = 0;
    constexpr int CH = D / 8;   // 16-byte chunks per row
    for (int i = tid; i < MQT * CH; i += 128) {
        const int r = i / CH, c = (i % CH) * 8;
        bf16* dst = Qs + r * LD + c;
        if (q0 + r < Sq) cp_async16(dst, Q + static_cast<size_t>(q0 + r) * a.ldq + c);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < n_rows * CH; i += 128) {
        const int r = i / CH, c = (i % CH) * 8;
        const int key = t_begin + r;
        bf16* dk = Ks + r * LD + c;
        bf16* dv = Vs + r * LD + c;
        if (key < Skv) {
            cp_async16(dk, K + static_cast<size_t>(key) * a.ldk + c);
            cp_async16(dv, V + static_cast<size_t>(key) * a.ldv + c);
        } else {
            *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();   // s_any_key initialised
    if (a.mask_kind == MASK_KEY_VECTOR) {
        const float* km = a.key_mask + static_cast<size_t>(b) * Skv;
        int any = 0;
        for (int k = tid; k < n_rows; k += 128) {
            const float mval = (t_begin + k < Skv) ? km[t_begin + k] : 0.f;
            Ms[k] = mval;
            any |= (mval > 0.5f);
        }
        if (any) atomicOr(&s_any_key, 1);
    }
    cp_async_wait_all();
    __syncthreads();
    const bool key_all = (a.mask_kind == MASK_KEY_VECTOR) && (s_any_key == 0);   // every key masked: uniform shift only

    // ---- main loop
    const float scale_log2 = a.scale * 1.4426950408889634f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

    const bf16* q_base = Qs + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * LD + 8 * (lane >> 4);
    for (int t0 = t_begin; t0 < t_end; t0 += MKT) {
        const bf16* k_tile = Ks + static_cast<size_t>(t0 - t_begin) * LD;
        const bf16* v_tile = Vs + static_cast<size_t>(t0 - t_begin) * LD;
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
        // scale (folded into the exp2 argument) + mask
        float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = t0 + nb * 8 + 2 * t + (e & 1);
                const int r = e >> 1;
                bool ok = key < Skv;
                if (a.mask_kind == MASK_KEY_VECTOR) ok = ok && (key_all || Ms[key - t_begin] > 0.5f);
                else ok = ok && ((key >= lo[r] && key < hi[r]) || key == self[r]);
                const float v = ok ? s[nb][e] * scale_log2 : -INFINITY;
                s[nb][e] = v;
                tmax[r] = fmaxf(tmax[r], v);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
            const float m_new = fmaxf(m_run[r], tmax[r]);
            corr[r] = (m_new == -INFINITY) ? 1.f : fast_exp2(m_run[r] - m_new);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
        const float mb0 = (m_run[0] == -INFINITY) ? 0.f : m_run[0];
        const float mb1 = (m_run[1] == -INFINITY) ? 0.f : m_run[1];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = fast_exp2(s[nb][0] - mb0);   // exp2(-inf) = 0 for masked entries
            s[nb][1] = fast_exp2(s[nb][1] - mb0);
            s[nb][2] = fast_exp2(s[nb][2] - mb1);
            s[nb][3] = fast_exp2(s[nb][3] - mb1);
            l_run[0] += s[nb][0] + s[nb][1];
            l_run[1] += s[nb][2] + s[nb][3];
        }
#pragma unroll
        for (int i = 0; i < D / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        // O += P V
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
    // ---- finalize: row sums across the quad, normalise, store bf16
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
        if (row0 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row0) * a.ldo + col) = pack2<FP16>(o[i][0] * inv0, o[i][1] * inv0);
        if (row1 < Sq) *reinterpret_cast<uint32_t*>(O + static_cast<size_t>(row1) * a.ldo + col) = pack2<FP16>(o[i][2] * inv1, o[i][3] * inv1);
    }
}

template <int D, bool FP16>
int launch_mma(const AttnArgs& a, cudaStream_t stream) {
    const int kv_rows_max = ((a.Skv + MKT - 1) / MKT) * MKT;
    const size_t smem = sizeof(bf16) * (MQT + 2 * static_cast<size_t>(kv_rows_max)) * (D + PADE) + sizeof(float) * kv_rows_max;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        UNIMM_CUDA_CHECK(cudaFuncSetAttribute(attn_mma_kernel<D, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    dim3 grid((a.Sq + MQT - 1) / MQT, a.heads, a.B);
    attn_mma_kernel<D, FP16><<<grid, 128, smem, stream>>>(a, kv_rows_max);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace

int attention_mma_lp(const AttnArgs& a, cudaStream_t stream) {
    UNIMM_CHECK(a.B > 0 && a.B <= 65535 && a.heads > 0 && a.Sq > 0 && a.Skv > 0 && a.Skv <= 256, "attention: bad problem size");
    UNIMM_CHECK(a.D == 64 || a.D == 128, "attention: head dim must be 64 or 128");
    UNIMM_CHECK((a.ldq % 8) == 0 && (a.ldk % 8) == 0 && (a.ldv % 8) == 0 && (a.ldo % 2) == 0, "attention: rows must be 16-byte aligned");
    UNIMM_CHECK(a.mask_kind == MASK_KEY_VECTOR ? a.key_mask != nullptr : a.desc != nullptr, "attention: mask operand missing");
    if (a.lp_kind == LP_FP16) return a.D == 64 ? launch_mma<64, true>(a, stream) : launch_mma<128, true>(a, stream);
    return a.D == 64 ? launch_mma<64, false>(a, stream) : launch_mma<128, false>(a, stream);
}

}  // namespace unimm
