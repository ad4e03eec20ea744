// Ranking metrics on the GPU: replaces the Python double loops of the reference's
// utils/visdial_metrics.py — scores_to_ranks (:21-39), SparseGTMetrics (:52-90: R@1/5/10, mean rank, MRR) and
// NDCG (:122-176) — for the [units, 100] score tensor the scoring path produces.
//
// One CTA per row of n_opt scores.  rank_j = 1 + #{i : s_i > s_j} + #{i < j : s_i == s_j}: the ranking of a
// stable descending sort.  The reference sorts with an unstable sort, so on exact ties its order is arbitrary;
// ties are counted and returned so the caller can report them instead of silently comparing.
#include "common.cuh"
#include "kernels.h"

namespace unimm {
namespace {

// sums[0]=rows  [1]=#rank<=1  [2]=#rank<=5  [3]=#rank<=10  [4]=sum rank  [5]=sum 1/rank  [6]=sum ndcg  [7]=#ndcg rows  [8]=#tied pairs
__global__ void __launch_bounds__(128)
rank_metrics_kernel(const float* __restrict__ scores, int n_opt, const int* __restrict__ gt_index,
                    const float* __restrict__ relevance, int* __restrict__ ranks, double* __restrict__ sums) {
    extern __shared__ float sm[];
    float* s = sm;                         // [n_opt]
    float* rel = s + n_opt;                // [n_opt]
    int* rk = reinterpret_cast<int*>(rel + n_opt);   // [n_opt]
    __shared__ float red[8];
    __shared__ int ired[4];
    const int row = blockIdx.x, tid = threadIdx.x;
    for (int j = tid; j < n_opt; j += blockDim.x) {
        s[j] = scores[static_cast<size_t>(row) * n_opt + j];
        rel[j] = relevance ? relevance[static_cast<size_t>(row) * n_opt + j] : 0.f;
    }
    __syncthreads();
    int ties = 0, k_rel = 0;
    for (int j = tid; j < n_opt; j += blockDim.x) {
        const float sj = s[j];
        int r = 1;
        for (int i = 0; i < n_opt; ++i) {
            const float si = s[i];
            r += (si > sj) || (si == sj && i < j);
            ties += (si == sj && i < j);
        }
        rk[j] = r;
        if (ranks) ranks[static_cast<size_t>(row) * n_opt + j] = r;
        k_rel += (rel[j] != 0.f);
    }
    // block reductions of (ties, k_rel)
    for (int o = 16; o > 0; o >>= 1) {
        ties += __shfl_xor_sync(0xffffffffu, ties, o);
        k_rel += __shfl_xor_sync(0xffffffffu, k_rel, o);
    }
    if (tid < 4) ired[tid] = 0;
    __syncthreads();
    if ((tid & 31) == 0) { atomicAdd(&ired[0], ties); atomicAdd(&ired[1], k_rel); }
    __syncthreads();
    const int k = ired[1];
    if (tid == 0) {
        atomicAdd(&sums[0], 1.0);
        if (ired[0]) atomicAdd(&sums[8], static_cast<double>(ired[0]));
        if (gt_index) {
            const int g = gt_index[row];
            const int r = rk[g];
            if (r <= 1) atomicAdd(&sums[1], 1.0);
            if (r <= 5) atomicAdd(&sums[2], 1.0);
            if (r <= 10) atomicAdd(&sums[3], 1.0);
            atomicAdd(&sums[4], static_cast<double>(r));
            atomicAdd(&sums[5], 1.0 / static_cast<double>(r));
        }
    }
    if (relevance != nullptr) {
        // DCG over the k best-ranked options, ideal DCG over the k most relevant ones (k = #non-zero relevance)
        float dcg = 0.f, ideal = 0.f;
        for (int j = tid; j < n_opt; j += blockDim.x) {
            if (rk[j] <= k) dcg += rel[j] / log2f(static_cast<float>(rk[j]) + 1.0f);
            int rr = 1;
            const float rj = rel[j];
            for (int i = 0; i < n_opt; ++i) rr += (rel[i] > rj) || (rel[i] == rj && i < j);
            if (rr <= k) ideal += rj / log2f(static_cast<float>(rr) + 1.0f);
        }
        dcg = warp_sum(dcg);
        ideal = warp_sum(ideal);
        if (tid < 8) red[tid] = 0.f;
        __syncthreads();
        if ((tid & 31) == 0) { atomicAdd(&red[0], dcg); atomicAdd(&red[1], ideal); }
        __syncthreads();
        if (tid == 0 && k > 0) {
            atomicAdd(&sums[6], static_cast<double>(red[0] / red[1]));
            atomicAdd(&sums[7], 1.0);
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// NeuralNDCG (transposed) of one slate per CTA — the dense-annotation fine-tuning objective on the NSP probabilities
// (reference utils/rank_loss.py:518-581 as called at dense_annotation_finetuning.py:288: deterministic NeuralSort :79-112,
// Sinkhorn scaling :55-78, powered relevancies, no padding, k = n, forward value only).
//   P_hat[k][i] = softmax_i((s_i * (n + 1 - 2 (k + 1)) - sum_j |s_i - s_j|) / tau)
//   Sinkhorn: <= max_iter rounds of (columns, then rows) normalisation with the reference's early exit
//   ndcg = sum_i (2^y_i - 1) * (sum_k P[k][i] / log2(k + 2)) / (idcg + 1e-8)
// The n x n matrix lives in shared memory; thread t owns column t in the column phase and row t in the row phase.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
neural_ndcg_kernel(const float* __restrict__ y_pred, const float* __restrict__ y_true, int n, float inv_tau, int max_iter, float tol,
                   float* __restrict__ ndcg_out, float* __restrict__ idcg_out) {
    extern __shared__ float sm[];
    float* P = sm;                       // [n][n + 1] (padded rows: conflict-free column walks)
    const int ld = n + 1;
    float* s = P + static_cast<size_t>(n) * ld;   // [n]
    float* y = s + n;                    // [n]
    float* Bv = y + n;                   // [n]
    float* colsum = Bv + n;              // [n]
    __shared__ float red[4];
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
    // NeuralSort rows: thread k builds and normalises row k
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
    // Sinkhorn scaling
    for (int it = 0; it < max_iter; ++it) {
        for (int i = tid; i < n; i += blockDim.x) {             // dim = 1: every column divided by its sum over the rows
            float c = 0.f;
            for (int k = 0; k < n; ++k) c += P[k * ld + i];
            const float inv = 1.0f / fmaxf(c, 1e-8f);
            for (int k = 0; k < n; ++k) P[k * ld + i] *= inv;
        }
        __syncthreads();
        float worst = 0.f;
        for (int k = tid; k < n; k += blockDim.x) {             // dim = 2: every row divided by its sum
            float r = 0.f;
            for (int i = 0; i < n; ++i) r += P[k * ld + i];
            const float inv = 1.0f / fmaxf(r, 1e-8f);
            float r2 = 0.f;
            for (int i = 0; i < n; ++i) { const float v = P[k * ld + i] * inv; P[k * ld + i] = v; r2 += v; }
            worst = fmaxf(worst, fabsf(r2 - 1.0f));
        }
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {             // convergence test on both marginals (reference :70-71)
            float c = 0.f;
            for (int k = 0; k < n; ++k) c += P[k * ld + i];
            worst = fmaxf(worst, fabsf(c - 1.0f));
        }
        worst = warp_max(worst);
        if ((tid & 31) == 0) red[tid >> 5] = worst;
        __syncthreads();
        if (tid == 0) done = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) < tol;
        __syncthreads();
        if (done) break;
    }
    // expected discounts, gains, ideal DCG
    float num = 0.f, idcg = 0.f;
    for (int i = tid; i < n; i += blockDim.x) {
        float ed = 0.f;
        for (int k = 0; k < n; ++k) ed += P[k * ld + i] / log2f(static_cast<float>(k) + 2.0f);
        num += (exp2f(y[i]) - 1.0f) * ed;
        int r = 0;                                              // position of y_i in the descending order of y (ties by index)
        for (int j = 0; j < n; ++j) r += (y[j] > y[i]) || (y[j] == y[i] && j < i);
        idcg += (exp2f(y[i]) - 1.0f) / log2f(static_cast<float>(r) + 2.0f);
    }
    num = warp_sum(num);
    idcg = warp_sum(idcg);
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = num; colsum[tid >> 5] = idcg; }
    __syncthreads();
    if (tid == 0) {
        const float nsum = red[0] + red[1] + red[2] + red[3], isum = colsum[0] + colsum[1] + colsum[2] + colsum[3];
        idcg_out[row] = isum;
        ndcg_out[row] = isum == 0.f ? 0.f : nsum / (isum + 1e-8f);
    }
}

// ensemble of per-model option probabilities (reference val.py:152-161, evaluate.py:107-117): one warp per (row)
__global__ void __launch_bounds__(128)
ensemble_normalise_kernel(const float* __restrict__ probs, int n_models, int rows, int n_opt, float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    for (int j = lane; j < n_opt; j += 32) out[static_cast<size_t>(row) * n_opt + j] = 0.f;
    for (int m = 0; m < n_models; ++m) {
        const float* p = probs + (static_cast<size_t>(m) * rows + row) * n_opt;
        float lo = INFINITY, hi = -INFINITY;
        for (int j = lane; j < n_opt; j += 32) { lo = fminf(lo, p[j]); hi = fmaxf(hi, p[j]); }
        hi = warp_max(hi);
        lo = -warp_max(-lo);
        float sum = 0.f;
        for (int j = lane; j < n_opt; j += 32) sum += (p[j] - lo) / (hi - lo);
        sum = warp_sum(sum);
        for (int j = lane; j < n_opt; j += 32) out[static_cast<size_t>(row) * n_opt + j] += (p[j] - lo) / (hi - lo) / sum;
    }
}

}  // namespace

int rank_metrics(const float* scores, int rows, int n_opt, const int* gt_index, const float* relevance, int* ranks, double* sums,
                 cudaStream_t stream) {
    UNIMM_CHECK(rows > 0 && n_opt > 0 && n_opt <= 4096 && sums != nullptr, "rank metrics: bad arguments");
    const size_t smem = sizeof(float) * 3 * n_opt;
    rank_metrics_kernel<<<rows, 128, smem, stream>>>(scores, n_opt, gt_index, relevance, ranks, sums);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}

}  // namespace unimm

namespace unimm {
int neural_ndcg(const float* y_pred, const float* y_true, int rows, int n, float temperature, int max_iter, float tol, float* ndcg,
                float* idcg, cudaStream_t stream) {
    UNIMM_CHECK(rows > 0 && n > 0 && n <= 128 && temperature > 0.f && max_iter >= 0, "neural_ndcg: 1..128 options per slate");
    const size_t smem = sizeof(float) * (static_cast<size_t>(n) * (n + 1) + 4 * n);
    UNIMM_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(&neural_ndcg_kernel), smem));
    neural_ndcg_kernel<<<rows, 128, smem, stream>>>(y_pred, y_true, n, 1.0f / temperature, max_iter, tol, ndcg, idcg);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}
int ensemble_normalise(const float* probs, int n_models, int rows, int n_opt, float* out, cudaStream_t stream) {
    UNIMM_CHECK(n_models > 0 && rows > 0 && n_opt > 0, "ensemble_normalise: empty input");
    ensemble_normalise_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(probs, n_models, rows, n_opt, out);
    UNIMM_LAUNCH_CHECK(1);
    return 0;
}
}  // namespace unimm
