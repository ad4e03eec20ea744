// Internal launcher declarations shared by the .cu translation units of libunimm_b200.so.
// Every launcher returns 0 on success and records a message via set_error() otherwise.
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace unimm {

// Epilogue description shared by the tcgen05 and the SIMT GEMMs.  Plain mode: out = act(acc + bias)
// (+ residual), written as fp32 and/or bf16.  LSE mode (partials != nullptr, tcgen05 only): per row and
// per 256-wide column tile the pair (max, sum exp(x - max)) of x = acc + bias, and the logit at
// labels[row]; nothing else is written.
struct GemmEpilogue {
    const float* bias = nullptr;      // [N]
    const float* residual = nullptr;  // [M, ldr] fp32, added after the activation
    int ldr = 0;
    float* out_f32 = nullptr;
    int ldo_f32 = 0;
    bf16* out_bf16 = nullptr;
    int ldo_bf16 = 0;
    int act = ACT_NONE;
    int debug_mode = 0;            // microbenchmark only: 1 = row-per-thread stores, 2 = no stores, 3 = no epilogue work
    int lp_kind = LP_BF16;         // encoding of the 16-bit operands and of out_bf16 (LP_BF16 / LP_FP16)
    int in_kind = -1;              // encoding of the A / W operands when it differs from out_bf16's (-1: lp_kind) — the bf16 mode reads
                                   // LayerNorm outputs as fp16 and writes Q | K | V / GELU outputs as bf16
    int split3 = 0;                // fp32-class mode: A [M, 2K] and W [N, 2K] are fp16 hi | lo planes (LP_HILO); three passes per k-block
    bool out_hilo = false;         // write out_bf16 as fp16 hi | lo planes instead of one 16-bit value: hi at column c, lo at column
    int hilo_off = 0;              //   hilo_off + c of the same row (0 = N: an [M, 2N] matrix)
    bool w_perm16 = false;         // W rows are in fragment order (permute_weight_rows mode 1): enables the smem-free 16-bit epilogue
    float alpha = 1.f;             // plain fp32-output paths: out = act(alpha * acc + bias) (+ residual)
    const float* alpha_ptr = nullptr;   // the same factor read from device memory (wins; lets a scale computed on the device stay there)
    // LSE-mode variant "dz" (LM-head backward, heads.cu: lm_head_backward): instead of the partials, write
    //   dz[row, col] = coef[row] * (exp(x - lse[row]) - [col == labels[row]])     x = acc + bias
    // as 16-bit values to dz [M, ldz] (columns < dz_cols) and, transposed, to dzT [N, ldzt]
    bf16* dz = nullptr;
    int ldz = 0, dz_cols = 0;
    bf16* dzT = nullptr;
    int ldzt = 0;
    const float* lse = nullptr;    // [M]
    const float* coef = nullptr;   // [M]
    const int* labels = nullptr;   // LSE mode: [M]
    float2* partials = nullptr;    // LSE mode: [M, gemm_umma_lse_tiles(N)]
    float* label_logit = nullptr;  // LSE mode: [M]
    // operands given TRANSPOSED (MN-major for tcgen05: the non-contracted dimension is the contiguous one), so that the backward's
    // dgrad  dX = dY W  (W [N, K] as stored = B operand [K_out, N] MN-major)  and  wgrad  dW = dY^T X  (both operands as stored)
    // need no transposed copies:  a_mn: A is passed as [K, M] row-major (lda = its row stride);  b_mn: W is passed as [K, N] row-major.
    // With both set the contraction length K (= rows of both) may be any positive number (TMA zero-fills beyond it).
    bool a_mn = false, b_mn = false;
    // split the contraction over split_k CTAs per output tile, partial sums added to out_f32 with atomics (out_f32 must hold zeros or
    // the value to accumulate onto; fp32 output only, no bias / activation / residual): fills the machine when the output is small
    // and the contraction long — every wgrad of the path (N x K <= 3072 x 3072 outputs over 61 440 rows)
    int split_k = 1;
    // fp32-output paths: max |out| over the whole matrix, as float bits, by atomicMax (one per epilogue warp at the end of the kernel;
    // the caller zeroes it) — the next consumer's 16-bit operand scale without a pass over the matrix
    unsigned* amax_out = nullptr;
    // out_f32 receives the value BEFORE the activation (acc + bias) while out_bf16 receives act(acc + bias): the training forward keeps
    // the GELU's pre-activation for the backward and feeds the next GEMM in one epilogue (no residual in this mode)
    bool pre_act_f32 = false;
    // the same with the pre-activation stored as 16-bit values in lp_kind's encoding: out_f32 then POINTS AT a 16-bit [M, ldo_f32] matrix
    // (half the write traffic of the FFN-1 forward and half the read traffic of its backward's pass; gelu' moves by < 1e-3)
    bool pre_act_lp = false;
    // training forward: dropout on (acc + bias) BEFORE the residual is added (models/vilbert_dialog.py:424, :467, :553, :596, :746, :749);
    // element index = row * N + column
    DropArgs drop;
};

// C = A[M,K] · W[N,K]^T, bf16 operands, fp32 accumulation in TMEM (gemm_umma.cu).
// tile_n: 0 = auto, 128 or 256.  max_ctas: 0 = one persistent CTA per SM.
int gemm_umma_bf16(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmEpilogue& ep, int tile_n,
                   int max_ctas, cudaStream_t stream);
int gemm_umma_lse_tiles(int N);

// X = LayerNorm(A · W^T + bias + residual) * gamma + beta with the row statistics exchanged across a thread-block
// cluster (gemm_umma_ln.cu).  N must be 768 or 1024; residual may alias out_f32, residual_lp may alias out_lp (in-place residual stream).
struct GemmLnEpilogue {
    const float* bias = nullptr;      // [N]
    const float* residual = nullptr;  // [M, ldr] fp32 residual (precharged into the accumulator by the epilogue warps), or
    int ldr = 0;
    const bf16* residual_lp = nullptr;  // [M, ldr_lp] 16-bit residual (added by the tensor core as identity k-blocks); wins if set
    int ldr_lp = 0;
    const float* gamma = nullptr;     // [N]
    const float* beta = nullptr;      // [N]
    float* out_f32 = nullptr;         // [M, ldo_f32] fp32 master copy (optional)
    int ldo_f32 = 0;
    bf16* out_lp = nullptr;           // [M, ldo_lp] 16-bit GEMM-operand copy (optional)
    int ldo_lp = 0;
    int lp_kind = LP_BF16;            // encoding of out_lp (and of residual_lp)
    int in_kind = -1;                 // encoding of the A / W operands when it differs (-1: lp_kind)
    bool a_multicast = true;          // the cluster's CTAs share the activation box by TMA multicast (UNIMM_LN_MULTICAST=0 disables)
};
// W must be the row-permuted copy produced by permute_weight_rows_ln (see gemm_umma_ln.cu: the permutation makes each
// thread's tcgen05.ld fragment 8 consecutive output columns).
int permute_weight_rows_ln(const bf16* W, bf16* Wp, int N, int K, cudaStream_t stream);
// mode 0: order for the LayerNorm-fused kernel (a thread's fragment = output columns 4a..4a+3 and 16+4a..16+4a+3);
// mode 1: order for the 16-bit-output epilogue of gemm_umma_bf16 (fragment = output columns 8a..8a+7)
int permute_weight_rows(const bf16* W, bf16* Wp, int N, int K, int mode, cudaStream_t stream);
bool gemm_umma_ln_supported(int N, int K, const GemmLnEpilogue& ep);
int gemm_umma_ln(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmLnEpilogue& ep, cudaStream_t stream);

// Same contraction in fp32 on the CUDA cores (gemm_simt.cu) — the fp32 parity mode.
int gemm_simt_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                  cudaStream_t stream);

// ------------------------------------------------------------------------------------------ rowwise.cu
// text embeddings: word + position + (type | type-extension) gather, sum, LayerNorm (ref :326-356)
int embed_text_ln(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int rows, int H, int vocab,
                  int max_pos, int type_vocab, int type_ext, const float* word_emb, const float* pos_emb,
                  const float* type_emb, const float* type_ext_emb, const float* gamma, const float* beta, float* out_f32,
                  bf16* out_bf16, int lp_kind, int* err_flag, cudaStream_t stream);
int embed_text_ln_i32(const int32_t* ids, const int32_t* type_ids, const int32_t* pos_ids, int rows, int H, int vocab,
                      int max_pos, int type_vocab, int type_ext, const float* word_emb, const float* pos_emb,
                      const float* type_emb, const float* type_ext_emb, const float* gamma, const float* beta, float* out_f32,
                      bf16* out_bf16, int lp_kind, int* err_flag, cudaStream_t stream);
// y = LayerNorm(x) * gamma + beta, eps 1e-12 inside the sqrt, biased variance; x may alias y_f32.
int layernorm_rows(const float* x, int ldx, int rows, int H, const float* gamma, const float* beta, float* y_f32,
                   bf16* y_bf16, int lp_kind, cudaStream_t stream);
// backward of layernorm_rows (dx may not alias dy; dgamma / dbeta [H] are zeroed here) and of the erf GELU
// amax_out (optional, device): receives max |dx| as float bits (zeroed here), so that the consumer that turns dx into a 16-bit operand
// (linear backward) needs no pass of its own over it
int layernorm_backward(const float* dy, const float* x, int rows, int H, const float* gamma, float* dx, float* dgamma, float* dbeta,
                       cudaStream_t stream, float* amax_out = nullptr);
int gelu_backward(const float* dy, const float* x, size_t n, float* dx, cudaStream_t stream, float* amax_out = nullptr);
// image location term: out[r, :] = loc[idx(r), 0:5] · Wloc[H,5]^T + bloc  (fp32, K = 5)
int image_loc_embed(const float* loc, const int* feat_index, int B, int R, int H, const float* Wloc, const float* bloc,
                    float* out, cudaStream_t stream);
// gather image features into the (optionally bf16) GEMM operand: dst[b*R + r] = feat[feat_index[b]*R + r]
int gather_features(const float* feat, const int* feat_index, int B, int R, int F, float* dst_f32, bf16* dst_bf16,
                    int lp_kind, cudaStream_t stream);
int gather_rows(const float* src_f32, const bf16* src_bf16, const int* rows, int n, int H, float* dst_f32, bf16* dst_bf16,
                cudaStream_t stream);
int cast_f32_to_lp(const float* src, bf16* dst, size_t n, int lp_kind, cudaStream_t stream);
// fp32 [rows, K] -> fp16 hi | lo planes [rows, 2K] (LP_HILO, the operand layout of the fp32-class tcgen05 GEMM)
// (row_idx: optional gather — output row r reads source row row_idx[r])
int split_f32_to_hilo(const float* x, int ldx, int rows, int K, bf16* out, cudaStream_t stream, const int* row_idx = nullptr);
// *d_flag = 1 iff the two fp32 buffers differ in any bit
int buffers_differ(const float* a, const float* b, size_t n, int* d_flag, cudaStream_t stream);
int cast_lp_to_f32(const bf16* src, float* dst, size_t n, int lp_kind, cudaStream_t stream);
int gather_labels(const int64_t* labels, const int* rows, int n, int* out, cudaStream_t stream);
int expand_key_mask(const float* mask, const int* index, int B, int R, float* out, cudaStream_t stream);

// ------------------------------------------------------------------------------------------ attention.cu
enum MaskKind : int {
    MASK_TEXT_SELF = 0,  // per-row interval (+ self) from the sequence descriptor (gen / dis text self-attention)
    MASK_KEY_VECTOR = 1, // float key mask [B, Skv] (image padding mask): text->image and image self-attention
    MASK_CO_INTERVAL = 2 // column interval from the descriptor (image->text co-attention)
};
struct AttnArgs {
    const void* q; int ldq;   // [B*Sq, ...] rows; head h at column h*D
    const void* k; int ldk;   // [B*Skv, ...]
    const void* v; int ldv;
    void* o; int ldo;         // [B*Sq, heads*D]
    int B, heads, D, Sq, Skv;
    int mask_kind;
    const SeqDesc* desc;      // [B]
    const float* key_mask;    // [B, Skv] (MASK_KEY_VECTOR)
    float scale;              // 1/sqrt(D)
    int lp_kind;              // encoding of 16-bit q/k/v/o (ignored by the fp32 kernel)
    DropArgs drop;            // dropout on the attention probabilities (training; mma.sync kernel and its backward); element index
                              //   ((b * heads + h) * Sq + q) * Skv + k
    float* lse = nullptr;     // optional [B, heads, Sq]: log sum_k exp(scale * q.k) over the allowed keys (mma.sync kernel only; saved
                              //   for attention_backward_lp)
};
int attention_simt_f32(const AttnArgs& a, cudaStream_t stream);
int attention_simt_lp(const AttnArgs& a, cudaStream_t stream);
int attention_mma_lp(const AttnArgs& a, cudaStream_t stream);

// ------------------------------------------------------------------------------------------ attention_jobs.cu
// Prefix-shared layout: attention as a list of jobs over packed rows (see attention_jobs.cu).
struct AttnJobsArgs {
    const void* q; int ldq;
    const void* k; int ldk;
    const void* v; int ldv;
    void* o; int ldo;
    int heads, D;
    const int* jobs;       // [n_jobs, 8] = q_start, q_len, kv_start, kv_len, win, mask_row, -, -
    int n_jobs;
    int max_q_len;         // longest q_len among the jobs (grid sizing)
    int kv_cap;            // staged rows for the shared range, multiple of 64, >= every kv_len
    int win_cap;           // staged rows for the per-CTA window, multiple of 64 (0 when no job has win)
    const int* row_iv;     // [rows, 4] = lo, hi, self, - (packed row indices) for rows of win jobs
    const float* key_mask; // [units, key_mask_ld] for jobs with mask_row >= 0
    int key_mask_ld;
    float scale;
    int lp_kind;
    int n_rows = 0;        // rows of the q (and, for self-attention, k / v) matrices: TMA tensor extent of the tcgen05 kernels
    int n_kv_rows = 0;     // rows of the k / v matrices when they differ from q's (cross attention)
    const SeqDesc* desc = nullptr;   // dense layout (attention_dense_umma): one descriptor per sequence
    int seq_len = 0;                 //   rows per sequence (<= 256)
    // fp32-class mode (attention_jobs_split): q / k / v / o are fp16 hi planes, the lo plane of the same row lies lo_off_*
    // elements further
    int lo_off_q = 0, lo_off_k = 0, lo_off_v = 0, lo_off_o = 0;
};
// attention_split.cu: the same job semantics with hi | lo fp16 planes and three mma.sync passes per product (fp32-class mode)
int attention_jobs_split(const AttnJobsArgs& a, cudaStream_t stream);
// job lists / per-row intervals of the DENSE [B, S] layout from the descriptors: jobs_text / jobs_i2t / jobs_img [B, 8], row_iv [B*S, 4]
int build_dense_jobs(const SeqDesc* desc, int B, int S, int R, int* jobs_text, int* jobs_i2t, int* jobs_img, int* row_iv,
                     cudaStream_t stream);
int attention_jobs(const AttnJobsArgs& a, bool fp32, cudaStream_t stream);
// jobs that all have win = 1 (candidate rows over context + own rows), D = 64, 16-bit: persistent double-buffered kernel.
// halo = (longest candidate's row count - 1): rows of a candidate lie within +-halo of any of its rows.
int attention_candidates(const AttnJobsArgs& a, int halo, cudaStream_t stream);
// the same jobs on tcgen05 (attention_umma.cu): context part on the 5th-gen tensor cores with S / P / O in TMEM, own-candidate
// part with mma.sync on a TMA-staged window; needs halo <= 16 and a.n_rows
bool attention_candidates_umma_supported(const AttnJobsArgs& a, int halo);
int attention_candidates_umma(const AttnJobsArgs& a, int halo, cudaStream_t stream);
// dense [B, seq_len] text self-attention under the descriptor masks on tcgen05 (a.desc, a.seq_len, a.n_jobs = B, a.n_rows = B * seq_len)
bool attention_dense_umma_supported(const AttnJobsArgs& a);
int attention_dense_umma(const AttnJobsArgs& a, cudaStream_t stream);
// window-free jobs over <= 64 keys with D = 128 (text -> image co-attention) on tcgen05; needs n_rows and n_kv_rows
bool attention_cross_umma_supported(const AttnJobsArgs& a);
int attention_cross_umma(const AttnJobsArgs& a, cudaStream_t stream);

// ------------------------------------------------------------------------------------------ heads.cu
// NSP head on the pooled vectors (ref :1062-1070 with fusion 'mul'): nsp[b, c] = sum_j pooled_t[b, j] * pooled_v[b, j] * Wn[c, j] + bn[c]
int nsp_from_pooled(const float* pooled_t, const float* pooled_v, int n, int Hb, const float* Wn, const float* bn, float* nsp_logits,
                    cudaStream_t stream);
// per-candidate sum of the compact per-row log-probs: rows [off[c], off[c+1])
int segment_sum(const float* vals, const int* off, int C, float* out, cudaStream_t stream);
// fp32 LM head tail: per row log-softmax pick from materialised logits (fp32 mode only)
int lse_from_logits(const float* logits, int ld, int rows, int V, const int* labels, float* logp, float* ul,
                    cudaStream_t stream);
// LM-head backward building blocks (heads.cu)
// 16-bit transpose: dst[c, r] = src[r, c] for r < rows, c < cols (dst rows ldt wide; the caller zeroes any padding)
int transpose_16(const bf16* src, int lds, int rows, int cols, bf16* dst, int ldt, cudaStream_t stream);
// per labelled row: coef = scale * w (likelihood row, w > 0) or -scale * p / (1 - p) (unlikelihood row, w == -1; 0 where 1 - p was
// clamped at 1e-6), p = exp(logp)
int lm_loss_coef(const float* logp, const float* weight, int n, float scale, float* coef, cudaStream_t stream);
int row_sums_16(const bf16* x, int ld, int rows, int cols, int lp_kind, float alpha, float* out, cudaStream_t stream);
// dgrad / wgrad helpers (unimm_k_linear_backward): out2[0] = 2^k with max|x| * 2^k in [2^9, 2^10) (1 when want_scale == 0 or x == 0),
// out2[1] = 1 / out2[0];  y16 = lp(x * out2[0]);  colsum[j] = sum_i x[i, j]
int amax_scale(const float* x, size_t n, int want_scale, float* out2, cudaStream_t stream, const float* known_amax = nullptr, float headroom = 1.f);
// y = lp(v * scale[0]), colsum (optional, zeroed here) += v, with v = x or x * gelu'(gelu_t) (heads.cu: cast_colsum_kernel)
// gelu_t_kind: -1 = gelu_t is fp32; LP_BF16 / LP_FP16 = gelu_t points at 16-bit values of that encoding
int cast_colsum_lp(const float* x, int ldx, const float* gelu_t, int ldt, int rows, int cols, const float* scale, bf16* y, int ldy, int lp_kind,
                   float* colsum, cudaStream_t stream, DropArgs drop = DropArgs(), int gelu_t_kind = -1);
int cast_scaled_lp(const float* x, int ldx, int rows, int cols, const float* scale, bf16* y, int ldy, int lp_kind, cudaStream_t stream);
int column_sums_f32(const float* x, int ldx, int rows, int cols, float* out, cudaStream_t stream);
// tcgen05 LM head tail: merge the per-tile (max, sum) partials
int lse_merge(const float2* partials, int tiles, int rows, float* lse, cudaStream_t stream);
int label_scores(const bf16* h, int ldh, const bf16* E, int lde, const float* bias, const int* uidx, const int* labels, const float* lse,
                 int n, int K, int lp_kind, float* logp, float* ul, cudaStream_t stream);
int label_scores_f32(const float* h, int ldh, const float* E, int lde, const float* bias, const int* uidx, const int* labels, const float* lse,
                     int n, int K, float* logp, float* ul, cudaStream_t stream);
int lse_from_partials(const float2* partials, int tiles, const float* label_logit, int rows, float* logp, float* ul,
                      cudaStream_t stream);
// scatter the compact per-row results to dense [B,S] (zero elsewhere) and sum per sequence (val_lm.py:131-136)
int scatter_scores(const float* logp, const float* ul, const int* flat_rows, int n, int B, int S, float* token_logp,
                   float* token_ul, float* seq_score, cudaStream_t stream);
// likelihood / unlikelihood loss (ref :1577-1595); out[0] = loss
int lm_ul_loss(const float* logp, const float* ul, const int* flat_rows, int n, const int64_t* lm_weight, int BS,
               float* out, cudaStream_t stream);
// mean CE over rows when lm_weight is None (ref :1601-1604)
int lm_ce_loss(const float* logp, int n, float* out, cudaStream_t stream);
// weighted NSP cross entropy (ref :1605-1621)
int nsp_ce_loss(const float* nsp_logits, const int64_t* labels, int B, const float* nsp_weight, float* out,
                cudaStream_t stream);
// masked image KL (ref :1569-1574)
int image_kl_loss(const float* v_logits, int ld, const float* target, const int64_t* image_label, int rows, int C, float* out,
                  cudaStream_t stream);
// GPU ranking metrics (metrics.cu); sums is a zero-initialised double[9]
int rank_metrics(const float* scores, int rows, int n_opt, const int* gt_index, const float* relevance, int* ranks, double* sums,
                 cudaStream_t stream);
// NeuralNDCG-transposed forward value per slate (utils/rank_loss.py:518-581 as used by dense_annotation_finetuning.py:288); n <= 128
int neural_ndcg(const float* y_pred, const float* y_true, int rows, int n, float temperature, int max_iter, float tol, float* ndcg,
                float* idcg, cudaStream_t stream);
// ensemble of per-model option probabilities (val.py:152-161): probs [models, rows, n_opt] -> out [rows, n_opt]
int ensemble_normalise(const float* probs, int n_models, int rows, int n_opt, float* out, cudaStream_t stream);
// regenerate the dense masks from descriptors and compare with the caller's dense tensors (boundary check)
int verify_masks(const SeqDesc* desc, int B, int S, int R, const void* txt_mask, int txt_elem_bytes, int txt_is_2d,
                 const int64_t* co_mask, int* mismatch_flag, cudaStream_t stream);

}  // namespace unimm
