// Shared device/host helpers for the unimm_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace unimm {

// ---------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns an int and leaves a message in a thread-local string
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
const char* get_error();

#define UNIMM_CUDA_CHECK(expr)                                                                      \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::unimm::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));                          \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)

#define UNIMM_CHECK(cond, msg)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::unimm::set_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +             \
                               std::to_string(__LINE__));                                          \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

// after a kernel launch: count it (bench.py reports gpu_launches) and surface launch errors
void count_launches(int n);
#define UNIMM_LAUNCH_CHECK(n)                   \
    do {                                        \
        ::unimm::count_launches(n);             \
        UNIMM_CUDA_CHECK(cudaGetLastError());   \
    } while (0)

// Programmatic dependent launch (opt-in: UNIMM_PDL=1): kernels that execute griddepcontrol.wait before their first global access
// are launched with this attribute, so that their prologue (CTA launch, barrier init, TMEM allocation, tensor-map prefetch) overlaps
// the tail of the kernel before them.  Parity-green, but measured +-0.5 % on the bench step (the step is power-bound: idle gaps
// between kernels are not lost time, they buy clock), so it stays off by default.
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("UNIMM_PDL"); v = e ? (atoi(e) != 0) : 0; }
    return v != 0;
}
inline void add_pdl_attr(cudaLaunchAttribute* attr, unsigned* n) {
    if (!pdl_enabled()) return;
    attr[*n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[*n].val.programmaticStreamSerializationAllowed = 1;
    ++*n;
}

#define UNIMM_TRY(expr)            \
    do {                           \
        int _rc = (expr);          \
        if (_rc != 0) return _rc;  \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: raise it once per (device, kernel) — a process may
// hold engines on several devices (one handle per device) — and only upwards.  Thread-safe.
int ensure_dynamic_smem(const void* kernel, size_t bytes);

typedef __nv_bfloat16 bf16;

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 2^x on the SFU (ex2.approx.ftz: ~2 ulp, exact 0 for -inf)
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exact (erf) GELU, reference models/vilbert_dialog.py:115-121
__device__ __forceinline__ float gelu_erf(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_GELU_TANH = 3 /* fragment epilogue only: 1-SFU tanh form */,
                 ACT_GELU_ERF = 4 /* erff form of the fp32-class mode */ };

// erf-GELU for the tensor-core epilogues, where the accurate erff (~30 issue slots per element) makes the
// FFN-1 epilogue slower than its K=768 main loop.  GELU(x) = x * Phi(x) with Phi(x) = 1 / (1 + 2^(x * R(x^2))):
// R is a degree-4 minimax fit (in x^2) of -log2(e) * logit(Phi(x)) / x, weighted by the sensitivity of the
// result; max |gelu_fast - gelu_erf| = 3.5e-6 over all x evaluated in fp32 (R < 0 everywhere and -> -inf, so
// the tails saturate to x and -0 without a clamp).  Two SFU ops (ex2, rcp) and 8 FP32 ops per element — and on
// sm_100 the FP32 part runs two elements per instruction (fma.rn.f32x2 -> FFMA2), i.e. 6 issue slots per element
// against ~20 for a rational erf.  Far below the 16-bit rounding of the stored result; the fp32 parity mode keeps
// erff (gelu_erf).
namespace f32x2 {
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t dup(float a) { return pack(a, a); }
}  // namespace f32x2
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kGeluC0 = -2.302042459894489f, kGeluC1 = -0.1052317053540061f, kGeluC2 = 0.0003627706199861018f,
                kGeluC3 = 8.778429282804149e-05f, kGeluC4 = -3.2039596426976067e-06f;
__device__ __forceinline__ float gelu_fast(float x) {
    const float x2 = x * x;
    float r = fmaf(x2, kGeluC4, kGeluC3);
    r = fmaf(r, x2, kGeluC2);
    r = fmaf(r, x2, kGeluC1);
    r = fmaf(r, x2, kGeluC0);
    return x * fast_rcp(1.0f + fast_ex2(r * x));
}
// two elements at once on the packed-fp32 pipe
__device__ __forceinline__ void gelu_fast2(float& a, float& b) {
    using namespace f32x2;
    const uint64_t x = pack(a, b);
    const uint64_t x2 = mul(x, x);
    uint64_t r = fma(x2, dup(kGeluC4), dup(kGeluC3));
    r = fma(r, x2, dup(kGeluC2));
    r = fma(r, x2, dup(kGeluC1));
    r = fma(r, x2, dup(kGeluC0));
    float t0, t1;
    unpack(mul(r, x), t0, t1);
    float d0, d1;
    unpack(add(pack(fast_ex2(t0), fast_ex2(t1)), dup(1.0f)), d0, d1);
    unpack(mul(x, pack(fast_rcp(d0), fast_rcp(d1))), a, b);
}
// One-SFU-op variant: GELU(x) = h + h * tanh(x * T(x^2)), h = x / 2, T = the same fit rescaled (atanh(erf(x / sqrt 2)) / x).
// tanh.approx.f32 has a relative error of up to 2^-11, i.e. |error| <= |x| * 2.4e-4 — up to half of the fp16 rounding
// error of the stored result — in exchange for halving the SFU work of the FFN-1 epilogue (1 MUFU + 4 packed FP32 slots
// per element).  Selected per GEMM with ACT_GELU_TANH (engine: UNIMM_GELU_TANH=1); the default keeps the 3.5e-6 ex2/rcp form.
constexpr float kGeluT0 = 0.7978276471008666f, kGeluT1 = 0.03646983712264174f, kGeluT2 = -0.00012547381454677784f,
                kGeluT3 = -3.0459227792119586e-05f, kGeluT4 = 1.1119725660024423e-06f;
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_tanh2(float& a, float& b) {
    using namespace f32x2;
    const uint64_t x = pack(a, b);
    const uint64_t x2 = mul(x, x);
    uint64_t r = fma(x2, dup(kGeluT4), dup(kGeluT3));
    r = fma(r, x2, dup(kGeluT2));
    r = fma(r, x2, dup(kGeluT1));
    r = fma(r, x2, dup(kGeluT0));
    float t0, t1;
    unpack(mul(r, x), t0, t1);
    const uint64_t h = mul(x, dup(0.5f));
    unpack(fma(h, pack(fast_tanh(t0), fast_tanh(t1)), h), a, b);
}
#ifndef UNIMM_GELU_TANH
#define UNIMM_GELU_TANH 0
#endif
__device__ __forceinline__ void gelu_epi2(float& a, float& b) {
#if UNIMM_GELU_TANH
    gelu_tanh2(a, b);
#else
    gelu_fast2(a, b);
#endif
}
__device__ __forceinline__ float apply_act_fast(float x, int act) {
    if (act == ACT_GELU) return gelu_fast(x);
    if (act == ACT_RELU) return fmaxf(x, 0.f);
    return x;
}

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_GELU) return gelu_erf(x);
    if (act == ACT_RELU) return fmaxf(x, 0.f);
    return x;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

// The tensor-core path stores activations / weights in one of two 16-bit encodings chosen per engine.
// Buffers are typed bf16* throughout (storage only); `kind` says how the 16 bits are to be read.
//   LP_BF16  8-bit exponent, 7-bit mantissa
//   LP_FP16  5-bit exponent, 10-bit mantissa (8x finer rounding; values saturate at +-65504 instead of
//            overflowing — every tensor stored this way is LayerNorm-bounded, a GELU output, a QKV
//            projection or an attention context, far inside that range)
//   LP_HILO  (fp32-class mode) TWO fp16 planes per row: x = hi + lo with hi = fp16(x), lo = fp16(x - hi) — 22 mantissa bits;
//            a row of K values is stored as [hi_0 .. hi_{K-1} | lo_0 .. lo_{K-1}] (leading dimension >= 2K).  The tcgen05 GEMM
//            then accumulates a_lo*w_hi + a_hi*w_lo + a_hi*w_hi in its fp32 TMEM accumulator (gemm_umma.cu, split3)
enum LpKind : int { LP_BF16 = 0, LP_FP16 = 1, LP_HILO = 2 };
typedef __half fp16;

template <>
__device__ __forceinline__ float to_f32<fp16>(fp16 v) { return __half2float(v); }
template <>
__device__ __forceinline__ fp16 from_f32<fp16>(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }

__device__ __forceinline__ uint32_t pack_fp16x2(float lo, float hi) {
    uint32_t r;   // one F2FP.SATFINITE: round to nearest, clamp to +-65504 (first source operand -> upper half)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// two values -> their fp16 hi parts and the fp16 residuals
__device__ __forceinline__ void split_hilo2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_fp16x2(a, b);
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = pack_fp16x2(a - h.x, b - h.y);
}
__device__ __forceinline__ uint32_t pack_lp2(float lo, float hi, int kind) {
    return kind == LP_FP16 ? pack_fp16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ bf16 lp_from_f32(float v, int kind) {
    if (kind == LP_FP16) {
        const fp16 h = from_f32<fp16>(v);
        return *reinterpret_cast<const bf16*>(&h);
    }
    return __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// Counter-based dropout (training step): element i of a tensor at one dropout site of one forward is kept iff
// lowbias32(lowbias32(i ^ seed) + seed) >= p * 2^32 — a pure function of (seed, i), so the backward regenerates the mask instead of
// storing it (tests/torch_train_ops.py::keep_mask is the same function in numpy).  Kept values are scaled by 1 / (1 - p).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t idx, uint32_t thresh) { return lowbias32(lowbias32(idx ^ seed) + seed) >= thresh; }
struct DropArgs {
    uint32_t seed = 0, thresh = 0;   // thresh = 0: dropout off
    float scale = 1.f;               // 1 / (1 - p)
};
inline DropArgs make_drop(uint32_t seed, float p) {
    DropArgs d;
    if (p > 0.f) {
        const double t = static_cast<double>(p) * 4294967296.0;
        d.seed = seed; d.thresh = t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t); d.scale = 1.f / (1.f - p);
    }
    return d;
}

// ---------------------------------------------------------------------------------------------
// Sequence descriptor: the 4 integers that regenerate every dense attention mask of the reference
// (utils/data_utils.py:149-210 generative, :353-354 discriminative).
//   mode      0 = generative, 1 = discriminative
//   ctx       first row of the visible answer copy (= L - last_len); generative only
//   L         orig_length: number of tokens before the masked answer copy
//   last_len  answer length + 1 ([SEP]); T = L + last_len is the number of real rows (generative)
// ---------------------------------------------------------------------------------------------
struct SeqDesc {
    int mode, ctx, L, last_len;
};

// Allowed key interval [lo,hi) plus optional extra single column `self` (-1 = none) for text query row r.
// An empty result (lo >= hi and self < 0) marks a padding row: the reference adds -10000 to every
// column there, which leaves softmax over the raw scores — callers then attend to all S columns.
__device__ __forceinline__ void text_row_interval(const SeqDesc& d, int r, int S, int& lo, int& hi, int& self) {
    self = -1;
    if (d.mode == 1) {  // discriminative: [0,L) x [0,L)
        if (r < d.L) { lo = 0; hi = min(d.L, S); } else { lo = 0; hi = 0; }
        return;
    }
    const int T = d.L + d.last_len;
    if (r == 0) { lo = 0; hi = min(T, S); }
    else if (r < d.ctx) { lo = 1; hi = min(d.ctx, S); }
    else if (r < d.L) { lo = 1; hi = r + 1; }
    else if (r < T) { lo = 1; hi = max(1, min(r - d.last_len, S)); self = r; }
    else { lo = 0; hi = 0; }
}

// Allowed text columns for image queries (co-attention mask): gen [1,ctx), dis [0,L).
__device__ __forceinline__ void co_interval(const SeqDesc& d, int S, int& lo, int& hi) {
    if (d.mode == 1) { lo = 0; hi = min(d.L, S); } else { lo = 1; hi = min(d.ctx, S); }
    if (hi <= lo) { lo = 0; hi = 0; }
}

}  // namespace unimm
