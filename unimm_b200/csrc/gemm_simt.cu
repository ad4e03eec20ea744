// fp32 CUDA-core GEMM:  C[M,N] = act(A[M,K] · W[N,K]^T + bias) (+ residual)
//
// This is the "fp32 mode" of the hot path: same call sites as gemm_umma.cu, but every product and sum
// is an fp32 FMA, so per-candidate log-likelihoods agree with the reference's fp32 eval
// (val_lm.py has no autocast) to ~1e-5.  It is the accuracy leg, not the throughput leg: 128x128x16
// tiles, 256 threads, 8x8 register micro-tiles, register-staged double buffering.
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

constexpr int SBM = 128, SBN = 128, SBK = 16, SPAD = 4;

__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, int M, int N, int K,
                GemmEpilogue ep) {
    __shared__ __align__(16) float As[2][SBK][SBM + SPAD];
    __shared__ __align__(16) float Ws[2][SBK][SBN + SPAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    const int lrow = tid >> 2;        // 0..63
    const int lk = (tid & 3) * 4;     // 0,4,8,12

    float4 ra[2], rw[2];
    auto load_global = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = m0 + lrow + 64 * i;
            ra[i] = (r < M) ? *reinterpret_cast<const float4*>(A + static_cast<size_t>(r) * lda + k0 + lk)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            const int c = n0 + lrow + 64 * i;
            rw[i] = (c < N) ? *reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * ldw + k0 + lk)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i;
            As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
            Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y; Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = K / SBK;
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_global((kt + 1) * SBK);
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_smem(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c0 = n0 + h * 64 + tx * 4;
            float x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                float v = acc[i][h * 4 + j];
                if (c < N) {
                    if (ep.bias != nullptr) v += __ldg(ep.bias + c);
                    v = apply_act(v, ep.act);
                    if (ep.residual != nullptr) v += ep.residual[static_cast<size_t>(r) * ep.ldr + c];
                }
                x[j] = v;
            }
            if (ep.out_f32 != nullptr) {
                float* o = ep.out_f32 + static_cast<size_t>(r) * ep.ldo_f32 + c0;
                if (c0 + 3 < N && (ep.ldo_f32 & 3) == 0) {
                    *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + j < N) o[j] = x[j];
                }
            }
            if (ep.out_bf16 != nullptr) {
                bf16* o = ep.out_bf16 + static_cast<size_t>(r) * ep.ldo_bf16 + c0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c0 + j < N) o[j] = lp_from_f32(x[j], ep.lp_kind);
            }
        }
    }
}

}  // namespace

int gemm_simt_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                  cudaStream_t stream) {
    UNIMM_CHECK(M > 0 && N > 0 && K > 0 && K % SBK == 0, "simt gemm: K must be a positive multiple of 16");
    UNIMM_CHECK((lda & 3) == 0 && (ldw & 3) == 0, "simt gemm: leading dimensions must be multiples of 4");
    UNIMM_CHECK(ep.partials == nullptr, "simt gemm has no LSE epilogue");
    dim3 grid((N + SBN - 1) / SBN, (M + SBM - 1) / SBM);
    UNIMM_CHECK(grid.y <= 65535, "simt gemm: M too large for one launch");
    sgemm_nt_kernel<<<grid, 256, 0, stream>>>(A, lda, W, ldw, M, N, K, ep);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace unimm
