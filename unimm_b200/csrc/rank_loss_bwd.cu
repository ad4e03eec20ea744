// Gradient of the dense-annotation objective: neuralNDCG_transposed (reference utils/rank_loss.py:518-581 with
// the arguments of dense_annotation_finetuning.py:288) differentiated with respect to the predicted scores, i.e. what
// loss.backward() (dense_annotation_finetuning.py:296) sends into the NSP probabilities.
//
//   loss = -(1 / #valid) sum_r ndcg_r,   ndcg_r = sum_i (2^y_i - 1) ed_i / (idcg_r + eps),   ed_i = sum_k P[k, i] / log2(k + 2)
//   P = Sinkhorn^T(softmax_rows((s_i * scaling_k - B_i) / tau)),   B_i = sum_j |s_i - s_j|
//
// One CTA per slate.  The forward is replayed exactly as metrics.cu runs it (same arithmetic, same stopping iteration), keeping only
// the T x 2 scaling vectors of the Sinkhorn iterations; the backward then walks the iterations in reverse, un-normalising the matrix
// in place (P_before = P_after * sum) while it pulls the gradient through each normalisation
//   M' = M / c:   dM = (dM' - <dM', M'>) / c      per column / per row
// then through the row softmax of NeuralSort and the |s_i - s_j| sums.  The clamp max(sum, 1e-8) is treated as inactive (its sums
// are ~1 after the first iteration and >= e^-large > 1e-8 in the first); autograd's abs'(0) = 0 convention is kept (sign(0) = 0).
#include "../../include/unimm_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

__global__ void ndcg_count_valid_kernel(const float* __restrict__ y_true, int rows, int n, int* __restrict__ count) {
    // idcg != 0  <=>  some gain 2^y - 1 differs from 0 (gains are >= 0 for relevances >= 0)
    int c = 0;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        float idcg = 0.f;
        for (int i = 0; i < n; ++i) idcg += fabsf(exp2f(y_true[static_cast<size_t>(r) * n + i]) - 1.0f);
        c += idcg != 0.f;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void __launch_bounds__(128)
neural_ndcg_bwd_kernel(const float* __restrict__ y_pred, const float* __restrict__ y_true, int n, float inv_tau, int max_iter, float tol,
                       float grad_scale, const int* __restrict__ n_valid, float* __restrict__ d_pred, float* __restrict__ ndcg_out) {
    extern __shared__ float sm[];
    const int ld = n + 1;
    float* P = sm;                                   // [n][n + 1]
    float* G = P + static_cast<size_t>(n) * ld;      // [n][n + 1] gradient with respect to the current matrix
    float* s = G + static_cast<size_t>(n) * ld;      // [n]
    float* y = s + n;
    float* Bv = y + n;
    float* va = Bv + n;                              // scratch vectors
    float* vb = va + n;
    float* sc_col = vb + n;                          // [max_iter][n] 1 / column sum of iteration t
    float* sc_row = sc_col + static_cast<size_t>(max_iter) * n;   // [max_iter][n] 1 / row sum
    __shared__ float red[4], red2[4];
    __shared__ int done;
    const int row = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) {
        s[i] = y_pred[static_cast<size_t>(row) * n + i];
        y[i] = y_true[static_cast<size_t>(row) * n + i];
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        float b = 0.f;
        for (int j = 0; j < n; ++j) b += fabsf(s[i] - s[j]);
        Bv[i] = b;
    }
    __syncthreads();
    // ---- forward replay (metrics.cu: neural_ndcg_kernel)
    for (int k = tid; k < n; k += blockDim.x) {
        const float sc = static_cast<float>(n + 1 - 2 * (k + 1));
        float mx = -INFINITY;
        for (int i = 0; i < n; ++i) {
            const float z = (s[i] * sc - Bv[i]) * inv_tau;
            P[k * ld + i] = z;
            mx = fmaxf(mx, z);
        }
        float sum = 0.f;
        for (int i = 0; i < n; ++i) {
            const float e = expf(P[k * ld + i] - mx);
            P[k * ld + i] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
        for (int i = 0; i < n; ++i) P[k * ld + i] *= inv;
    }
    __syncthreads();
    int T = 0;
    for (int it = 0; it < max_iter; ++it) {
        for (int i = tid; i < n; i += blockDim.x) {
            float c = 0.f;
            for (int k = 0; k < n; ++k) c += P[k * ld + i];
            const float inv = 1.0f / fmaxf(c, 1e-8f);
            sc_col[it * n + i] = inv;
            for (int k = 0; k < n; ++k) P[k * ld + i] *= inv;
        }
        __syncthreads();
        float worst = 0.f;
        for (int k = tid; k < n; k += blockDim.x) {
            float r = 0.f;
            for (int i = 0; i < n; ++i) r += P[k * ld + i];
            const float inv = 1.0f / fmaxf(r, 1e-8f);
            sc_row[it * n + k] = inv;
            float r2 = 0.f;
            for (int i = 0; i < n; ++i) { const float v = P[k * ld + i] * inv; P[k * ld + i] = v; r2 += v; }
            worst = fmaxf(worst, fabsf(r2 - 1.0f));
        }
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            float c = 0.f;
            for (int k = 0; k < n; ++k) c += P[k * ld + i];
            worst = fmaxf(worst, fabsf(c - 1.0f));
        }
        worst = warp_max(worst);
        if ((tid & 31) == 0) red[tid >> 5] = worst;
        __syncthreads();
        if (tid == 0) done = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) < tol;
        __syncthreads();
        T = it + 1;
        if (done) break;
    }
    // ---- value, and the gradient with respect to the final matrix
    float num = 0.f, idcg = 0.f;
    for (int i = tid; i < n; i += blockDim.x) {
        float ed = 0.f;
        for (int k = 0; k < n; ++k) ed += P[k * ld + i] / log2f(static_cast<float>(k) + 2.0f);
        num += (exp2f(y[i]) - 1.0f) * ed;
        int r = 0;
        for (int j = 0; j < n; ++j) r += (y[j] > y[i]) || (y[j] == y[i] && j < i);
        idcg += (exp2f(y[i]) - 1.0f) / log2f(static_cast<float>(r) + 2.0f);
    }
    num = warp_sum(num);
    idcg = warp_sum(idcg);
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = num; red2[tid >> 5] = idcg; }
    __syncthreads();
    const float nsum = red[0] + red[1] + red[2] + red[3], isum = red2[0] + red2[1] + red2[2] + red2[3];
    if (tid == 0 && ndcg_out != nullptr) ndcg_out[row] = isum == 0.f ? 0.f : nsum / (isum + 1e-8f);
    const int nv = *n_valid;
    if (isum == 0.f || nv == 0) {                    // the slate is masked out of the mean (reference :573-579): no gradient
        for (int i = tid; i < n; i += blockDim.x) d_pred[static_cast<size_t>(row) * n + i] = 0.f;
        return;
    }
    const float w = -grad_scale / (static_cast<float>(nv) * (isum + 1e-8f));
    for (int k = tid; k < n; k += blockDim.x) {
        const float dk = w / log2f(static_cast<float>(k) + 2.0f);
        for (int i = 0; i < n; ++i) G[k * ld + i] = dk * (exp2f(y[i]) - 1.0f);
    }
    __syncthreads();
    // ---- Sinkhorn iterations in reverse
    for (int it = T - 1; it >= 0; --it) {
        for (int k = tid; k < n; k += blockDim.x) {                 // row normalisation
            const float inv = sc_row[it * n + k];
            float dot = 0.f;
            for (int i = 0; i < n; ++i) dot += G[k * ld + i] * P[k * ld + i];
            const float r = 1.0f / inv;
            for (int i = 0; i < n; ++i) {
                G[k * ld + i] = (G[k * ld + i] - dot) * inv;
                P[k * ld + i] *= r;
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {                 // column normalisation
            const float inv = sc_col[it * n + i];
            float dot = 0.f;
            for (int k = 0; k < n; ++k) dot += G[k * ld + i] * P[k * ld + i];
            const float c = 1.0f / inv;
            for (int k = 0; k < n; ++k) {
                G[k * ld + i] = (G[k * ld + i] - dot) * inv;
                P[k * ld + i] *= c;
            }
        }
        __syncthreads();
    }
    // ---- row softmax of NeuralSort: dz = P_hat (G - <G, P_hat>)
    for (int k = tid; k < n; k += blockDim.x) {
        float dot = 0.f;
        for (int i = 0; i < n; ++i) dot += G[k * ld + i] * P[k * ld + i];
        for (int i = 0; i < n; ++i) G[k * ld + i] = P[k * ld + i] * (G[k * ld + i] - dot);
    }
    __syncthreads();
    // z[k, i] = (s_i scaling_k - B_i) / tau
    for (int i = tid; i < n; i += blockDim.x) {
        float a = 0.f, g = 0.f;
        for (int k = 0; k < n; ++k) {
            const float dz = G[k * ld + i];
            a += dz * static_cast<float>(n + 1 - 2 * (k + 1));
            g += dz;
        }
        va[i] = a * inv_tau;       // direct term
        vb[i] = -g * inv_tau;      // d loss / d B_i
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        float d = va[i];
        for (int j = 0; j < n; ++j) {
            const float diff = s[i] - s[j];
            const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            d += (vb[i] + vb[j]) * sg;
        }
        d_pred[static_cast<size_t>(row) * n + i] = d;
    }
}

// y_pred = softmax(nsp logits)[:, 0] (dense_annotation_finetuning.py:267-287) and its backward onto the logits (accumulating)
__global__ void nsp_prob0_kernel(const float* __restrict__ logits, int B, float* __restrict__ p0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float a = logits[2 * i], b = logits[2 * i + 1], mx = fmaxf(a, b);
    const float ea = expf(a - mx), eb = expf(b - mx);
    p0[i] = ea / (ea + eb);
}
__global__ void nsp_prob0_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dp0, int B, float* __restrict__ dlogits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float a = logits[2 * i], b = logits[2 * i + 1], mx = fmaxf(a, b);
    const float ea = expf(a - mx), eb = expf(b - mx);
    const float p = ea / (ea + eb), g = dp0[i] * p * (1.0f - p);
    dlogits[2 * i] += g;
    dlogits[2 * i + 1] -= g;
}

}  // namespace
}  // namespace unimm

using namespace unimm;

extern "C" {

int unimm_neural_ndcg_backward(const float* d_y_pred, const float* d_y_true, int rows, int n_opt, float temperature, int max_iter, float tol,
                               float grad_scale, float* d_dpred, float* d_ndcg, int32_t* d_scratch_count, void* stream) {
    UNIMM_CHECK(d_y_pred && d_y_true && d_dpred && d_scratch_count, "null argument");
    UNIMM_CHECK(rows > 0 && n_opt > 0 && n_opt <= 128 && temperature > 0.f && max_iter >= 0, "neural_ndcg backward: 1..128 options per slate");
    const size_t smem = sizeof(float) * (2 * static_cast<size_t>(n_opt) * (n_opt + 1) + 5 * n_opt + 2 * static_cast<size_t>(max_iter) * n_opt);
    UNIMM_CHECK(smem <= 227 * 1024, "neural_ndcg backward: max_iter x options does not fit the shared memory of an SM");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&neural_ndcg_bwd_kernel), smem));
    UNIMM_CUDA_CHECK(cudaMemsetAsync(d_scratch_count, 0, sizeof(int), st));
    ndcg_count_valid_kernel<<<1, 128, 0, st>>>(d_y_true, rows, n_opt, d_scratch_count);
    neural_ndcg_bwd_kernel<<<rows, 128, smem, st>>>(d_y_pred, d_y_true, n_opt, 1.0f / temperature, max_iter, tol, grad_scale, d_scratch_count,
                                                   d_dpred, d_ndcg);
    UNIMM_LAUNCH_CHECK(2);
    return 0;
}

int unimm_t_nsp_prob0(const float* d_logits, int B, float* d_p0, void* stream) {
    UNIMM_CHECK(d_logits && d_p0 && B > 0, "bad argument");
    nsp_prob0_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_logits, B, d_p0);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

int unimm_t_nsp_prob0_backward(const float* d_logits, const float* d_dp0, int B, float* d_dlogits_accum, void* stream) {
    UNIMM_CHECK(d_logits && d_dp0 && d_dlogits_accum && B > 0, "bad argument");
    nsp_prob0_bwd_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_logits, d_dp0, B, d_dlogits_accum);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // extern "C"
