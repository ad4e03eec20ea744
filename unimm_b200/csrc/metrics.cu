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
