// Device helpers shared by the tensor-core attention kernels (attention_mma.cu, attention_jobs.cu):
// cp.async staging, ldmatrix / mma.sync.m16n8k16 wrappers, the ex2 fast path and the 64-bit tile masks.
#pragma once

#include "common.cuh"

namespace unimm {
namespace attn {

constexpr int MKT = 64;   // keys per inner tile
constexpr int PADE = 8;   // 16-bit elements of row padding: 16 B shifts successive rows by 4 banks (ldmatrix conflict-free)

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
// 2^x for x <= 0 on the FMA / integer pipes (no SFU): x = n + f with n = round(x), f in [-0.5, 0.5]; 2^f by a degree-4 minimax
// polynomial (max relative error 7.6e-6 — far below the 16-bit rounding of the stored probability), 2^n by an exponent-field add.
// Two lanes at once on the packed-fp32 pipe.  An EXPERIMENT (UNIMM_ATTN_DBG=16): half of the probabilities of the tcgen05 attention's
// pass 2 computed this way instead of on the 16-per-clock ex2 unit.  Parity-green, but measured 1 % slower on the bench step (the
// softmax warps are issue- / latency-bound, the ~9 extra issue slots per element pair cost more than the SFU queue they relieve),
// so it is off by default; anything below 2^-125 comes out as ~0.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
    using namespace f32x2;
    constexpr float kMagic = 12582912.f;          // 1.5 * 2^23: adding it leaves round(x) in the low mantissa bits
    x0 = fmaxf(x0, -125.f);
    x1 = fmaxf(x1, -125.f);
    const float t0 = x0 + kMagic, t1 = x1 + kMagic;
    const uint64_t f = pack(x0 - (t0 - kMagic), x1 - (t1 - kMagic));
    uint64_t r = fma(dup(0.009278289042413235f), f, dup(0.05586502328515053f));
    r = fma(r, f, dup(0.2403089553117752f));
    r = fma(r, f, dup(0.6931295394897461f));
    r = fma(r, f, dup(0.9999978542327881f));
    float q0, q1;
    unpack(r, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
__device__ __forceinline__ unsigned long long low_bits(int n) {   // n in [0,64]
    return n >= 64 ? ~0ull : ((1ull << n) - 1ull);
}
// bits [t0, t0+64) of the allowed set  [lo,hi) U {self}
__device__ __forceinline__ unsigned long long tile_mask(int lo, int hi, int self, int t0) {
    const int a = min(max(lo - t0, 0), 64), b = min(max(hi - t0, 0), 64);
    unsigned long long m = (b > a) ? (low_bits(b) & ~low_bits(a)) : 0ull;
    if (self >= t0 && self < t0 + 64) m |= 1ull << (self - t0);
    return m;
}


// second interval variant: allowed = [lo1,hi1) U [lo2,hi2) U {self}
__device__ __forceinline__ unsigned long long tile_mask2(int lo1, int hi1, int lo2, int hi2, int self, int t0) {
    return tile_mask(lo1, hi1, -1, t0) | tile_mask(lo2, hi2, self, t0);
}

}  // namespace attn
}  // namespace unimm
