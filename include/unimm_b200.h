/*
 * unimm_b200 — C ABI of the B200 (sm_100a) generative-scoring hot path of UniMM-UL.
 *
 * This library replaces, behind the reference's own boundary, the body of
 *   VisualDialogEncoder.forward                      (reference models/visual_dialog_encoder.py:18-50)
 *   -> BertForMultiModalPreTraining.forward           (reference models/vilbert_dialog.py:1519-1626)
 *   -> BertModel.forward / BertEncoder.forward        (reference models/vilbert_dialog.py:1359-1472, :817-937)
 * plus the scoring rule of val_lm.py:131-137.  The reference has no FFI of its own (it is eager
 * PyTorch); the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md, which is
 * exactly what unimm_b200/_lib.py does.
 *
 * Conventions: every function returns 0 on success, non-zero on failure; unimm_last_error() then
 * returns a thread-local message.  No exceptions cross the ABI.  All buffers are caller-owned.
 * Pointers named d_* are device pointers on the engine's device, h_* are host pointers.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  One engine per device; an engine is
 * not re-entrant (one forward at a time), distinct engines are independent.
 */
#ifndef UNIMM_B200_H
#define UNIMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIMM_ABI_VERSION 3

typedef struct unimm_engine unimm_engine_t;

/* compute precision of the projections / attention */
enum { UNIMM_PREC_FP32 = 0, /* CUDA-core fp32: the <=1e-4 parity mode                              */
       UNIMM_PREC_BF16 = 1, /* tcgen05 bf16 x bf16 -> fp32 (TMEM)                                  */
       UNIMM_PREC_FP16 = 2  /* tcgen05 fp16 x fp16 -> fp32 (TMEM): same rate, 8x finer rounding    */ };
/* 16-bit encodings of the single-kernel entry points below */
enum { UNIMM_LP_BF16 = 0, UNIMM_LP_FP16 = 1 };

/* Mirrors the fields of the reference's BertConfig that the hot path reads
 * (models/vilbert_dialog.py:131-247, config/bert_base_6layer_6conect.json). */
typedef struct {
    int32_t vocab_size, hidden_size, num_hidden_layers, num_attention_heads, intermediate_size;
    int32_t max_position_embeddings, type_vocab_size;
    int32_t v_feature_size, v_target_size, v_hidden_size, v_num_hidden_layers, v_num_attention_heads, v_intermediate_size;
    int32_t bi_hidden_size, bi_num_attention_heads;
    int32_t num_connections;
    int32_t v_biattention_id[16];
    int32_t t_biattention_id[16];
    int32_t seq_len;      /* S: text positions per sequence (reference max_seq_len, 256)  */
    int32_t num_regions;  /* R: image regions per sequence incl. the global one (37)      */
} unimm_config_t;

/* The 4 integers per sequence that regenerate the reference's dense masks
 * (utils/data_utils.py:149-210 generative, :353-354 discriminative):
 * mode 0 = generative, 1 = discriminative; ctx = L - last_len; L = orig_length; last_len = answer length + 1. */
typedef struct {
    int32_t mode, ctx, L, last_len;
} unimm_seq_desc_t;

/* One forward call = one chunk of B sequences (B <= max_sequences given at creation).
 * Argument meaning follows VisualDialogEncoder.forward (visual_dialog_encoder.py:18-20). */
typedef struct {
    int32_t B;
    const int64_t* d_input_ids;        /* [B,S]                                                         */
    const int64_t* d_token_type_ids;   /* [B,S]  segments                                               */
    const int64_t* d_position_ids;     /* [B,S]  token_position_ids                                     */
    const unimm_seq_desc_t* d_desc;    /* [B]    replaces attention_mask [B,S,S] + co_attention_mask     */
    const float* d_image_feat;         /* [U,R,v_feature_size]                                          */
    const float* d_image_loc;          /* [U,R,5]                                                       */
    const float* d_image_mask;         /* [U,R]  image_attention_mask, values in {0,1}                  */
    const int32_t* d_feat_index;       /* [B] sequence -> image slot u in [0,U); NULL = identity (U = B) */
    const int64_t* d_masked_lm_labels; /* [B,S]  -1 = ignore; may be NULL when n_lm_rows == 0           */
    const int32_t* d_lm_rows;          /* [n_lm_rows] sorted flat positions b*S+s with label != -1      */
    int32_t n_lm_rows;
    /* training targets: the loss branch runs iff labels, next_sentence_label and image_target are all
     * non-NULL (visual_dialog_encoder.py:28-29) */
    const int64_t* d_lm_weight;           /* [B,S] or NULL (plain masked-LM CE, vilbert_dialog.py:1601) */
    const int64_t* d_next_sentence_label; /* [B] or NULL                                                */
    const int64_t* d_image_label;         /* [B,R] or NULL                                              */
    const float* d_image_target;          /* [B,R,v_target_size] or NULL                                */
    const float* d_nsp_weight;            /* [2] or NULL (= [1,1])                                      */
} unimm_batch_t;

/* Every output pointer is optional (NULL = not wanted). */
typedef struct {
    float* d_seq_score;          /* [B]    sum of answer-token log-probs (val_lm.py:131-136)             */
    float* d_token_logp;         /* [B,S]  log p(label) at labelled positions, 0 elsewhere (= -nll)      */
    float* d_token_ul;           /* [B,S]  log(max(1-p(label),1e-6)) (vilbert_dialog.py:1587), 0 elsewhere */
    float* d_nsp_scores;         /* [B,2]  seq_relationship_score                                        */
    float* d_losses;             /* [8]    [0]=masked_lm_loss [1]=masked_img_loss [2]=nsp_loss, rest scratch */
    float* d_sequence_output_t;  /* [B,S,hidden_size] final text hidden states                           */
    float* d_sequence_output_v;  /* [B,R,v_hidden_size]                                                  */
    float* d_prediction_scores_t;/* [B,S,vocab] full logits — compatibility path only, 31 MB / sequence  */
} unimm_outputs_t;

const char* unimm_last_error(void);
int unimm_abi_version(void);

int unimm_create(const unimm_config_t* cfg, int device, int precision, int max_sequences, unimm_engine_t** out);
int unimm_destroy(unimm_engine_t* e);
/* Accepts the reference checkpoint keys (with or without the "bert_pretrained." prefix); h_data is fp32. */
int unimm_load_weight(unimm_engine_t* e, const char* name, const float* h_data, const int64_t* shape, int ndim);
/* Packs (QKV concatenation, bf16 casts) and verifies that every key the forward needs was loaded. */
int unimm_finalize_weights(unimm_engine_t* e);
int unimm_forward(unimm_engine_t* e, const unimm_batch_t* batch, const unimm_outputs_t* out, void* stream);
/* The reference's nn.Embedding raises on a token / position / token-type id outside its table (models/vilbert_dialog.py:334-350).
 * The asynchronous forwards clamp such an id and set a device flag; this call synchronises `stream`, returns non-zero if any
 * forward on it since the previous check saw one, and clears the flag.  The host-buffer entry points (unimm_score_host, unimm_score_packed_host) do this themselves. */
int unimm_check_ids(unimm_engine_t* e, void* stream);

/* ---- prefix-shared generative scoring --------------------------------------------------------------------
 * In generative mode the context rows [1,ctx) and the image rows are identical for the 100 candidates of a
 * dialog round (they attend only context + image; utils/data_utils.py:199-210), so a unit = (image, round)
 * stores them once and each candidate contributes only its own rows [CLS, A_0..A_{last-1}, B_0..B_{last-1}].
 * Text rows are packed: all units' context rows, then all candidates' rows.  Attention is described by job
 * lists over packed rows (8 int32 per job: q_start, q_len, kv_start, kv_len, win, mask_row, 0, 0) and, for
 * candidate rows, by row_iv[r] = (lo, hi, self, 0): the packed rows of its own candidate it may attend
 * besides the whole context.  unimm_b200/packing.py builds all of this.
 * Scores-only batches (no_cls_rows = 1) leave out the two candidate rows no labelled position attends ([CLS] and
 * A_{last-1}); the engine then also skips what only the pooled NSP logit reads (the image stream after the last
 * connection) and runs the row-wise half of the last text layer on the labelled rows alone. */
typedef struct {
    int32_t n_units, n_cands, n_text_rows;       /* U, C, M                                                   */
    const int32_t* d_input_ids;                  /* [M]                                                        */
    const int32_t* d_token_type_ids;             /* [M]                                                        */
    const int32_t* d_position_ids;               /* [M]                                                        */
    const int32_t* d_row_iv;                     /* [M,4]                                                      */
    const float* d_image_feat;                   /* [U,R,v_feature_size]                                       */
    const float* d_image_loc;                    /* [U,R,5]                                                    */
    const float* d_image_mask;                   /* [U,R]                                                      */
    const int32_t* d_jobs_text_self;             /* [n,8] first the context-self jobs, then the candidate jobs */
    int32_t n_jobs_text_self, max_q_text_self;   /*       (win = 1)                                            */
    int32_t n_jobs_text_ctx;                     /* how many of them are context-self jobs                     */
    int32_t cand_halo;                           /* longest candidate's row count - 1                          */
    const int32_t* d_jobs_t2i;                   /* [n,8] text rows of a unit over its image rows              */
    int32_t n_jobs_t2i, max_q_t2i;
    const int32_t* d_jobs_i2t;                   /* [U,8] image rows over the unit's context rows              */
    int32_t n_jobs_i2t;
    const int32_t* d_jobs_img_self;              /* [U,8]                                                      */
    int32_t n_jobs_img_self;
    int32_t kv_cap_text;                         /* multiple of 64, >= longest context, <= 256                  */
    int32_t win_cap;                             /* multiple of 64, >= 128 + 2*(longest candidate row count)    */
    const int32_t* d_lm_rows;                    /* [n_lm] packed rows of the masked copy, grouped by candidate */
    const int32_t* d_lm_labels;                  /* [n_lm]                                                      */
    int32_t n_lm_rows;
    const int32_t* d_cand_lm_off;                /* [C+1] offsets into lm rows                                  */
    const int32_t* d_cand_cls_row;               /* [C]   packed row of each candidate's [CLS]                  */
    const int32_t* d_cand_img_row;               /* [C]   unit * R                                              */
    double pairs_text_self, pairs_i2t;           /* sum over jobs of q_len * keys (profiling only)             */
    int32_t n_shared_rows;                       /* rows [0, n_shared_rows) are the units' context rows: the only text rows whose
                                                  * co-attention keys / values anything reads (0 = unknown: project all rows) */
    int32_t no_cls_rows;                         /* 1: packed for the sequence scores only — the candidates carry no [CLS] row (and no
                                                  * A_{last-1} row), d_cand_cls_row is unused and NSP scores cannot be requested   */
    /* optional: d_lm_rows lists some row more than once (the unit-wide B_0 row of a scores-only batch, once per candidate with that
     * candidate's label).  d_lm_urows [n_lm_unique] = the distinct labelled rows, d_lm_uidx [n_lm] = index of entry i's row in it;
     * the 16-bit modes then run the LM head once per distinct row.  n_lm_unique = 0: not given. */
    const int32_t* d_lm_urows;
    const int32_t* d_lm_uidx;
    int32_t n_lm_unique;
    /* optional: one feature / box / mask block per IMAGE instead of per unit (the 10 rounds of an image share it, val_lm.py:84-93
     * copies it x1000): d_image_feat / d_image_loc / d_image_mask then hold n_images blocks, d_unit_image [U] names each unit's
     * block and the mask_row field of the t2i / img_self jobs is an image index.  NULL: one block per unit (n_images ignored). */
    const int32_t* d_unit_image;
    int32_t n_images;
} unimm_packed_batch_t;
/* outputs (each optional): seq_score [C], nsp_scores [C,2], token_logp [n_lm] */
int unimm_forward_packed(unimm_engine_t* e, const unimm_packed_batch_t* batch, float* d_seq_score, float* d_nsp_scores,
                         float* d_token_logp, void* stream);
/* Same with every pointer of `hb` (and the outputs) in HOST memory: H2D + forward + D2H + sync inside the call. */
int unimm_score_packed_host(unimm_engine_t* e, const unimm_packed_batch_t* hb, float* h_seq_score, float* h_nsp_scores, void* stream);
/* The same in two halves, so that a sweep keeps the device busy across steps: submit enqueues H2D + forward + D2H of the batch on
 * `stream` using staging slot 0 or 1 and returns at once (the host arrays of `hb` and the result buffers must stay untouched until
 * the matching wait); wait blocks until that slot's scores have landed and reports an out-of-range id like unimm_check_ids.
 * Submit step i + 1 into the other slot before waiting for step i. */
int unimm_submit_packed_host(unimm_engine_t* e, const unimm_packed_batch_t* hb, int slot, float* h_seq_score, float* h_nsp_scores, void* stream);
int unimm_wait_packed(unimm_engine_t* e, int slot);

/* ---- packing on the host, from the reference's own layout -------------------------------------------------
 * What val_lm.py:55-121 holds for one step BEFORE the dense masks exist: per image one int64 [rows, S] tensor per field
 * (dataloader_visdial.py:437-457; rows = rounds x options) and one feature block.  A unit = (image, round) is a row
 * range of its image's block.  unimm_packer_pack turns a step of such blocks into the prefix-shared layout above (row
 * gathers, per-row intervals, job lists, labelled-row lists) inside the packer's own (pinned, when a CUDA device is
 * present) host buffers; unimm_packer_batch returns the unimm_packed_batch_t whose pointers name those buffers — pass it
 * to unimm_score_packed_host.  Host code only: threads = worker threads over units (0 = 4).  A packer is not
 * re-entrant; two packers double-buffer a sweep (pack step i + 1 while the device scores step i).
 * Sequences truncated at S (utils/data_utils.py:205-209, :237-244: L + last_len > S) keep the rows that exist.
 * desc == NULL: (ctx, L, last_len) are derived from position_ids (the masked copy restarts at position ctx, :227). */
typedef struct unimm_packer unimm_packer_t;
typedef struct {
    int32_t rows;                        /* sequences in this block                                     */
    const int64_t* input_ids;            /* [rows,S]                                                     */
    const int64_t* token_type_ids;       /* [rows,S]                                                     */
    const int64_t* position_ids;         /* [rows,S]                                                     */
    const int64_t* masked_lm_labels;     /* [rows,S]  -1 = ignore                                        */
    const unimm_seq_desc_t* desc;        /* [rows] or NULL                                               */
    const float* image_feat;             /* [R,F]                                                        */
    const float* image_loc;              /* [R,5]                                                        */
    const float* image_mask;             /* [R]                                                          */
} unimm_image_block_t;
typedef struct {
    int32_t n_blocks;
    const unimm_image_block_t* blocks;
    int32_t n_units;
    const int32_t* unit_block;           /* [U] block (image) of every unit, non-decreasing              */
    const int32_t* unit_row0;            /* [U] first sequence of the unit inside its block              */
    const int32_t* unit_rows;            /* [U] candidates of the unit                                   */
    int32_t scores_only;                 /* 1: no [CLS] / trailing A rows (see no_cls_rows above)        */
    int32_t share_first_mask;            /* 1: one B_0 row per unit when the candidates' agree           */
    int32_t verify_shared;               /* 1: compare every candidate's context with candidate 0's      */
} unimm_flat_batch_t;
int unimm_packer_create(int seq_len, int num_regions, int feature_size, int pinned, unimm_packer_t** out);
int unimm_packer_destroy(unimm_packer_t* p);
int unimm_packer_pack(unimm_packer_t* p, const unimm_flat_batch_t* fb, int threads);
/* the batch of the last successful unimm_packer_pack (HOST pointers, valid until the next pack / destroy) */
int unimm_packer_batch(const unimm_packer_t* p, unimm_packed_batch_t* out);
/* descriptors the last pack used (given or derived), in unit order: [n_cands] */
int unimm_packer_desc(const unimm_packer_t* p, const unimm_seq_desc_t** out, int32_t* n);

/* Boundary helper: check that the caller's dense masks equal what the descriptors regenerate.
 * d_txt_mask: [B,S,S] elements of txt_elem_bytes (1 = bool/uint8, 8 = int64); d_co_mask: [B,R,S] int64 or NULL.
 * *d_mismatch (device int) is set to 1 on any difference. */
int unimm_verify_masks(const unimm_seq_desc_t* d_desc, int B, int S, int R, const void* d_txt_mask, int txt_elem_bytes,
                       const int64_t* d_co_mask, int* d_mismatch, void* stream);

/* Host-buffer scoring entry (the end-to-end path of bench.py): stages the inputs through pinned memory,
 * copies host->device, runs unimm_forward, copies seq_score [B] and nsp_scores [B,2] back, synchronises.
 * Host arrays have the shapes of unimm_batch_t; U = number of distinct image slots. */
typedef struct {
    int32_t B, U;
    const int64_t* h_input_ids;
    const int64_t* h_token_type_ids;
    const int64_t* h_position_ids;
    const int64_t* h_masked_lm_labels;
    const unimm_seq_desc_t* h_desc;
    const float* h_image_feat;
    const float* h_image_loc;
    const float* h_image_mask;
    const int32_t* h_feat_index; /* NULL = identity */
} unimm_host_batch_t;
int unimm_score_host(unimm_engine_t* e, const unimm_host_batch_t* hb, float* h_seq_score, float* h_nsp_scores, void* stream);

/* Per-kernel-class device timing for bench.py: between begin and end every GEMM / attention / LayerNorm /
 * LM-head launch of this engine is bracketed by CUDA events on its launch stream.  end() synchronises and
 * returns, for classes 0 = tcgen05 or fp32 GEMM, 1 = attention, 2 = LayerNorm rows, 3 = fused LM-head GEMM,
 * 4 = other: summed milliseconds, summed algorithmic work (FLOPs; bytes for class 2) and launch counts.
 * ncat must be >= 5. */
int unimm_profile_begin(unimm_engine_t* e);
int unimm_profile_end(unimm_engine_t* e, double* ms, double* work, int64_t* launches, int ncat);
/* Algorithmic HBM bytes per class of the region the last unimm_profile_end closed (class 0 only: operands and weights once,
 * every output and residual once; 0 for the classes that do not state them).  bench.py puts them next to the ncu DRAM traffic. */
int unimm_profile_bytes(unimm_engine_t* e, double* bytes, int ncat);

/* Dense-annotation objective on the NSP probabilities, forward values:
 * unimm_neural_ndcg: utils/rank_loss.py:518-581 (neuralNDCG_transposed: deterministic NeuralSort :79-112 + Sinkhorn scaling :55-78,
 *   powered relevancies, no padded entries, k = n_opt <= 128) as called at dense_annotation_finetuning.py:288.  d_y_pred / d_y_true
 *   [rows, n_opt]; out: d_ndcg [rows] (0 where the ideal DCG is 0), d_idcg [rows].  loss = -sum(ndcg) / #(idcg != 0).
 * unimm_ensemble_normalise: val.py:152-161 / evaluate.py:107-117: per model min-max over the options, sum-normalise, sum over
 *   models.  d_probs [models, rows, n_opt] -> d_out [rows, n_opt]. */
int unimm_neural_ndcg(const float* d_y_pred, const float* d_y_true, int rows, int n_opt, float temperature, int max_iter, float tol,
                      float* d_ndcg, float* d_idcg, void* stream);
int unimm_ensemble_normalise(const float* d_probs, int n_models, int rows, int n_opt, float* d_out, void* stream);
/* Gradient of the dense-annotation objective (utils/rank_loss.py:518-581 under dense_annotation_finetuning.py:296's
 * loss.backward()): d_dpred [rows, n_opt] = grad_scale * d neuralNDCG_transposed(y_pred, y_true) / d y_pred (the loss is -mean of the
 * per-slate NDCG over the slates whose ideal DCG is not 0); d_ndcg (optional) [rows] the per-slate values; d_scratch_count: one int32.
 * The Sinkhorn iterations are replayed and walked in reverse inside one CTA per slate. */
int unimm_neural_ndcg_backward(const float* d_y_pred, const float* d_y_true, int rows, int n_opt, float temperature, int max_iter, float tol,
                               float grad_scale, float* d_dpred, float* d_ndcg, int32_t* d_scratch_count, void* stream);
/* y_pred of that objective: p0 = softmax(NSP logits [B, 2])[:, 0] (dense_annotation_finetuning.py:267-287), and its backward ADDED onto
 * d_dlogits_accum [B, 2] */
int unimm_t_nsp_prob0(const float* d_logits, int B, float* d_p0, void* stream);
int unimm_t_nsp_prob0_backward(const float* d_logits, const float* d_dp0, int B, float* d_dlogits_accum, void* stream);

/* Ranking metrics of the reference's utils/visdial_metrics.py on the device: d_scores [rows, n_opt]; optional d_gt_index
 * [rows] (sparse metrics), d_relevance [rows, n_opt] (NDCG), d_ranks [rows, n_opt] out (1-based, stable on ties).
 * d_sums: 9 doubles, zeroed by the caller, accumulated: rows, #rank<=1, #rank<=5, #rank<=10, sum rank, sum 1/rank,
 * sum ndcg, #ndcg rows, #tied pairs. */
int unimm_rank_metrics(const float* d_scores, int rows, int n_opt, const int32_t* d_gt_index, const float* d_relevance,
                       int32_t* d_ranks, double* d_sums, void* stream);

/* counters for bench.py: kernels launched by this library since the last reset */
int64_t unimm_launch_count(void);
void unimm_reset_launch_count(void);

/* ---- single-kernel entry points (used by tests/ to check each kernel against the oracle) ----
 * unimm_k_lm_head_lp scratch: d_partials_scratch holds rows * 2*ceil(V/256) float2, d_label_logit_scratch rows floats. */
/* unimm_k_gemm_lp: act = 0 none, 1 GELU, 2 ReLU; act | 0x100: d_out_f32 receives the PRE-activation (acc + bias) and d_out_lp the activation
 * — the training forward keeps the GELU's input for the backward and feeds the next GEMM from one epilogue; act | 0x200: the same with
 * the pre-activation stored as 16-bit values of lp_kind's encoding (d_out_f32 then points at a 16-bit [M, ldo_f32] matrix, ldo_f32 % 4 == 0);
 * unimm_k_linear_backward_acc / _phase take such a matrix as d_gelu_t when called with lp_kind | 0x100. */
int unimm_k_gemm_lp(const void* d_A_lp, int lda, const void* d_W_lp, int ldw, int M, int N, int K, const float* d_bias,
                    const float* d_residual, int ldr, int act, float* d_out_f32, int ldo_f32, void* d_out_lp, int ldo_lp,
                    int tile_n, int max_ctas, int lp_kind, void* stream);
/* LayerNorm(A W^T + bias + residual) * gamma + beta in one cluster-fused tcgen05 kernel (N = 768 or 1024; replaces the
 * reference's dense -> "+ input_tensor" -> LayerNorm tails, models/vilbert_dialog.py:422-426, :465-469, :745-752).
 * The residual is either fp32 (d_residual, may alias d_out_f32) or, when d_residual_lp is non-NULL, the 16-bit activation
 * copy itself (may alias d_out_lp), which the kernel adds on the tensor core.  d_W_lp must be the row-permuted copy made by unimm_k_permute_w_ln (the engine
 * makes it once at weight-load time). */
int unimm_k_permute_w_ln(const void* d_W_lp, void* d_Wp_lp, int N, int K, void* stream);
/* mode 0 = the order above; mode 1 = the order of unimm_k_gemm_lp's 16-bit-output epilogue (pass lp_kind | 0x100 there to
 * say that d_W_lp is such a copy: each thread's TMEM fragment is then 8 consecutive output columns, stored without a
 * shared-memory transpose). */
int unimm_k_permute_w(const void* d_W_lp, void* d_Wp_lp, int N, int K, int mode, void* stream);
int unimm_k_gemm_ln_lp(const void* d_A_lp, int lda, const void* d_W_lp, int ldw, int M, int N, int K, const float* d_bias,
                       const float* d_residual, int ldr, const void* d_residual_lp, int ldr_lp, const float* d_gamma,
                       const float* d_beta, float* d_out_f32, int ldo_f32, void* d_out_lp, int ldo_lp, int lp_kind, void* stream);
int unimm_k_gemm_f32(const float* d_A, int lda, const float* d_W, int ldw, int M, int N, int K, const float* d_bias,
                     const float* d_residual, int ldr, int act, float* d_out_f32, int ldo_f32, void* stream);
int unimm_k_lm_head_lp(const void* d_H_lp, int ldh, const void* d_E_lp, int lde, int rows, int V, int K,
                       const float* d_bias, const int32_t* d_labels, float* d_partials_scratch, float* d_label_logit_scratch,
                       float* d_logp, float* d_ul, int lp_kind, void* stream);
/* Backward of the fused LM head + likelihood / unlikelihood loss (first piece; reference
 * models/vilbert_dialog.py:1577-1595 under train.py:453-463's backward): for rows with hidden states H [rows, K] (16-bit), tied
 * decoder E [V, K] (16-bit), bias [V], labels and token weights w (w > 0: likelihood row, loss -w log p; w == -1: unlikelihood row,
 * loss -log(max(1 - p, 1e-6))), the gradients of  grad_scale * sum_i loss_i :
 *   dH [rows, K], dE [V, K] (decoder / word-embedding weight), dbias [V]   (fp32; d_dbias and d_logp optional).
 * No [rows, V] fp32 logits: the vocabulary GEMM is recomputed twice on tcgen05 (online log-sum-exp, then dz = coef (softmax - onehot)
 * written as 16-bit operands in both orientations) and dH = dz E, dE = dz^T H are two more tcgen05 GEMMs.
 * d_scratch: unimm_k_lm_head_backward_scratch(rows, V, K) bytes. */
size_t unimm_k_lm_head_backward_scratch(int rows, int V, int K);
int unimm_k_lm_head_backward(const void* d_H_lp, int ldh, const void* d_E_lp, int lde, int rows, int V, int K, const float* d_bias,
                             const int32_t* d_labels, const float* d_weight, float grad_scale, float* d_dH, float* d_dE, float* d_dbias,
                             float* d_logp, void* d_scratch, size_t scratch_bytes, int lp_kind, void* stream);
/* dgrad / wgrad of one nn.Linear  y = x W^T + b  on tcgen05 (SURVEY.md 8f item 1 building block; every projection of the path is one):
 * dY fp32 [M, N] (leading dimension ldy), X 16-bit [M, K], W 16-bit [N, K]  ->  dX = dY W fp32 [M, K], dW = dY^T X fp32 [N, K],
 * db = column sums of dY fp32 [N]  (each output optional).  N % 64 == 0.  d_scratch: unimm_k_linear_backward_scratch(M, N, K) bytes. */
size_t unimm_k_linear_backward_scratch(int M, int N, int K);
int unimm_k_linear_backward(const float* d_dY, int ldy, const void* d_X_lp, int ldx, const void* d_W_lp, int ldw, int M, int N, int K,
                            float* d_dX, float* d_dW, float* d_db, void* d_scratch, size_t scratch_bytes, int lp_kind, void* stream);
int unimm_k_layernorm(const float* d_x, int ldx, int rows, int H, const float* d_gamma, const float* d_beta, float* d_y_f32,
                      void* d_y_lp, int lp_kind, void* stream);
/* Backward of BertLayerNorm (models/vilbert_dialog.py:270-279; eps inside the sqrt, biased variance) and of the erf GELU (:115-121),
 * fp32: dx [rows, H], dgamma / dbeta [H] from dy and the layer's INPUT x; dx = dy * gelu'(x) (in place allowed). */
int unimm_k_layernorm_backward(const float* d_dy, const float* d_x, int rows, int H, const float* d_gamma, float* d_dx, float* d_dgamma,
                               float* d_dbeta, void* stream);
int unimm_k_gelu_backward(const float* d_dy, const float* d_x, int64_t n, float* d_dx, void* stream);
int unimm_k_cast_lp(const float* d_src, void* d_dst_lp, int64_t n, int lp_kind, void* stream);
/* unimm_k_attention (dense [B, S] layout): elem_kind = 0: fp32 tensors, 1: bf16, 2: fp16.  impl: 0 = CUDA-core kernel, 1 = mma.sync kernel (16-bit),
 * 2 = tcgen05 / TMEM kernel (16-bit, text self-attention with descriptor masks, D = 64, S <= 256). */
/* Attention over PACKED rows as a job list (csrc/attention_jobs.cu): every job = (q_start, q_len, kv_start, kv_len, win, mask_row, -, -);
 * rows of win jobs additionally attend [lo, hi) U {self} from d_row_iv[row] = (lo, hi, self, -).  16-bit tensors only.
 * impl 0 = generic job kernel, 1 = persistent mma.sync candidate kernel, 2 = tcgen05 / TMEM candidate kernel
 * (csrc/attention_umma.cu; D = 64, halo <= 16; n_rows = rows of the q/k/v matrices).  Replaces models/vilbert_dialog.py:395-410
 * for the rows a candidate owns under the generative mask of utils/data_utils.py:199-210. */
int unimm_k_attention_jobs(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo, int n_rows,
                           int heads, int D, const int32_t* d_jobs, int n_jobs, int max_q_len, int kv_cap, int win_cap,
                           const int32_t* d_row_iv, int halo, int lp_kind, int impl, void* stream);
/* Window-free jobs over <= 64 keys whose K / V live in another matrix (text -> image co-attention, models/vilbert_dialog.py:681-698):
 * d_key_mask[mask_row] (job[5]) marks valid keys; no valid key = all keys, as the reference's additive mask.
 * impl 0 = generic job kernel, 2 = tcgen05 / TMEM kernel (D = 128). */
int unimm_k_attention_cross_jobs(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo,
                                 int n_q_rows, int n_kv_rows, int heads, int D, const int32_t* d_jobs, int n_jobs, int max_q_len,
                                 const float* d_key_mask, int key_mask_ld, int lp_kind, int impl, void* stream);
int unimm_k_attention(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo, int B,
                      int heads, int D, int Sq, int Skv, int mask_kind, const unimm_seq_desc_t* d_desc,
                      const float* d_key_mask, int elem_kind, int impl, void* stream);

/* ---- training step (reference train.py:445-463: forward + loss.backward() + optimizer.step()) ----
 * Kernel-level entry points; the layer schedule in reverse and the saved activations are host logic (unimm_b200/train_step.py), as
 * autograd is in the reference.  All gradients are fp32; every tensor-core operand is 16-bit (lp_kind 0 = bf16, 1 = fp16). */
/* same as unimm_k_linear_backward; accumulate_dx != 0: dX += dY W (the residual branch's gradient is already in d_dX); d_amax (optional):
 * max |dY| as float bits, left on the device by the kernel that produced dY (the *_amax entry points below, unimm_k_attention_backward) —
 * the 16-bit operand scale is then derived without a pass over dY; d_gelu_t (optional, fp32 [M, N]): the projection is followed by the erf
 * GELU and d_dY is the gradient with respect to the GELU's OUTPUT: gelu'(d_gelu_t) is applied in the pass that casts dY and sums its
 * columns (no separate GELU-backward kernel, no fp32 copy of the pre-activation gradient); d_dX_amax (optional): receives max |dX| as
 * float bits from the dgrad GEMM's epilogue; drop_p > 0: the projection's OUTPUT went through nn.Dropout (unimm_t_gemm_drop with the same
 * seed): the same keep-mask / (1 - p) is applied to dY in that pass */
int unimm_k_linear_backward_acc(const float* d_dY, int ldy, const void* d_X_lp, int ldx, const void* d_W_lp, int ldw, int M, int N, int K,
                                float* d_dX, int accumulate_dx, float* d_dW, float* d_db, const float* d_amax, const float* d_gelu_t,
                                float* d_dX_amax, uint32_t drop_seed, float drop_p, void* d_scratch, size_t scratch_bytes, int lp_kind,
                                void* stream);
/* unimm_k_linear_backward_acc in two halves, so that the caller can put the wgrad GEMM on a second stream beside the (HBM-bound) work that
 * follows the dgrad: phase 1 = the pass over dY (16-bit copy and scale into d_scratch, bias gradient) + dgrad; phase 2 = wgrad alone from
 * the d_scratch a phase-1 call of the same shape filled; phase 0 = both (unimm_k_linear_backward_acc).  The caller orders the two streams
 * (phase 2 after phase 1's pass; the scratch not reused before phase 2 has run) — autograd's stream bookkeeping in the reference. */
int unimm_k_linear_backward_phase(const float* d_dY, int ldy, const void* d_X_lp, int ldx, const void* d_W_lp, int ldw, int M, int N, int K,
                                  float* d_dX, int accumulate_dx, float* d_dW, float* d_db, const float* d_amax, const float* d_gelu_t,
                                  float* d_dX_amax, uint32_t drop_seed, float drop_p, void* d_scratch, size_t scratch_bytes, int lp_kind,
                                  int phase, void* stream);
/* unimm_k_layernorm_backward / unimm_k_gelu_backward that also leave max |dx| (float bits; zeroed first) in d_amax[0] */
int unimm_k_layernorm_backward_amax(const float* d_dy, const float* d_x, int rows, int H, const float* d_gamma, float* d_dx, float* d_dgamma,
                                    float* d_dbeta, float* d_amax, void* stream);
int unimm_k_gelu_backward_amax(const float* d_dy, const float* d_x, int64_t n, float* d_dx, float* d_amax, void* stream);
/* mma.sync attention of the dense [B, S] layout (unimm_k_attention impl 1) that also saves the row log-sum-exp d_lse [B, heads, Sq]
 * (natural log, softmax scale included) for the backward. */
int unimm_k_attention_lse(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo, int B, int heads,
                          int D, int Sq, int Skv, int mask_kind, const unimm_seq_desc_t* d_desc, const float* d_key_mask, int lp_kind,
                          float* d_lse, uint32_t drop_seed, float drop_p, void* stream);
/* Backward of softmax(Q K^T / sqrt(D) + mask) V (models/vilbert_dialog.py:395-410, :681-721) from the saved 16-bit q / k / v / o and
 * d_lse: d_dO fp32 contiguous [B*Sq, heads*D] -> d_dq [B*Sq, lddq], d_dk / d_dv [B*Skv, lddk / lddv] fp32 (head h at column h*D, so the
 * three can be the column blocks of one [rows, 3H] matrix).  P is recomputed tile by tile; no [B, heads, Sq, Skv] tensor, no atomics.
 * The masks are the forward's, regenerated from d_desc / d_key_mask.  d_amax_accum (optional): atomicMax of |dq|, |dk|, |dv| as float bits,
 * NOT zeroed by the call (the two co-attentions fill column blocks of the same two gradient matrices); d_dO_amax (optional): max |dO| as
 * float bits when its producer left it on the device. */
size_t unimm_k_attention_backward_scratch(int B, int heads, int D, int Sq);
int unimm_k_attention_backward(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, const void* d_o, int ldo,
                               const float* d_dO, const float* d_lse, int B, int heads, int D, int Sq, int Skv, int mask_kind,
                               const unimm_seq_desc_t* d_desc, const float* d_key_mask, int lp_kind, float* d_dq, int lddq, float* d_dk,
                               int lddk, float* d_dv, int lddv, float* d_amax_accum, const float* d_dO_amax, uint32_t drop_seed, float drop_p,
                               void* d_scratch, size_t scratch_bytes, void* stream);
/* nn.Dropout of the reference's training mode (models/vilbert_dialog.py:355, :405, :424, :467, :534, :553, :596, :693, :716, :746, :749,
 * :1065, :1491), counter-based: element i is kept iff lowbias32(lowbias32(i ^ seed) + seed) >= p 2^32 and then scaled by 1 / (1 - p), so the
 * backward regenerates every mask from (seed, i).  unimm_t_dropout: y = dropout(x) as fp32 and / or 16-bit (forward; applied to a gradient
 * in place it is the backward).  unimm_t_gemm_drop: out = dropout(A W^T + bias) + residual, the mask applied in the GEMM epilogue
 * (element index row * N + column).  unimm_k_attention_lse / _backward take (drop_seed, drop_p) for the probabilities' dropout
 * (index ((b heads + h) Sq + q) Skv + k); drop_p = 0 everywhere = the reference in eval mode. */
int unimm_t_dropout(const float* d_x, int64_t n, uint32_t seed, float p, float* d_y_f32, void* d_y_lp, int lp_kind, void* stream);
int unimm_t_gemm_drop(const void* d_A_lp, int lda, const void* d_W_lp, int ldw, int M, int N, int K, const float* d_bias, const float* d_residual,
                      int ldr, uint32_t drop_seed, float drop_p, float* d_out_f32, int ldo_f32, int lp_kind, void* stream);
/* text embeddings without the LayerNorm (its input is what the backward needs): word + position + (type | type-extension)
 * (models/vilbert_dialog.py:334-352) -> d_out fp32 [rows, H]; and the scatter-add of that sum's gradient into the four tables
 * (accumulating: the word table's gradient also receives the tied decoder's). */
int unimm_t_embed_text_sum(const int64_t* d_ids, const int64_t* d_type_ids, const int64_t* d_pos_ids, int rows, int H, int vocab, int max_pos,
                           int type_vocab, int type_ext, const float* d_word, const float* d_pos, const float* d_type, const float* d_type_ext,
                           float* d_out, int* d_err_flag, void* stream);
int unimm_t_embed_text_backward(const float* d_dsum, const int64_t* d_ids, const int64_t* d_type_ids, const int64_t* d_pos_ids, int rows, int H,
                                int type_vocab, float* d_dword, float* d_dpos, float* d_dtype, float* d_dtype_ext, void* stream);
/* g = erf-GELU(t) written as the 16-bit operand of the next GEMM and / or as fp32 (t, the pre-activation, stays for unimm_k_gelu_backward) */
int unimm_t_gelu(const float* d_t, int64_t n, float* d_g_f32, void* d_g_lp, int lp_kind, void* stream);
/* value of the likelihood / unlikelihood loss (models/vilbert_dialog.py:1577-1595) from the per-row log p that unimm_k_lm_head_backward
 * returns: d_out[0] = scale * (sum_{w > 0} -w log p + sum_{w == -1} -log(max(1 - p, 1e-6))) */
int unimm_t_lm_ul_value(const float* d_logp, const float* d_weight, int n, float scale, float* d_out, void* stream);
/* fp32 element-wise: op 0: out = a + b, 1: out = a * b, 2: out = a * [b > 0], 3: out = alpha * a, 4: out = a + alpha * b (out may alias a / b) */
int unimm_t_ew(int op, int64_t n, const float* d_a, const float* d_b, float* d_out, float alpha, void* stream);
int unimm_t_gather_rows(const float* d_src, int lds, const int32_t* d_idx, int n, int H, float* d_dst, void* stream);
int unimm_t_scatter_add_rows(const float* d_src, const int32_t* d_idx, int n, int H, float* d_dst, int ldd, void* stream);
/* weighted NSP cross entropy (models/vilbert_dialog.py:1605-1621) and its gradient grad_scale * d loss / d logits [B, 2] */
int unimm_t_nsp_ce(const float* d_logits, const int64_t* d_labels, int B, const float* d_nsp_weight, float grad_scale, float* d_loss,
                   float* d_dlogits, void* stream);
/* masked image KL (:1569-1574) and its gradient [rows, ldd] (columns >= C and unselected rows are zeroed); the target distribution of row r
 * is row d_target_row[r] of d_target [., C] (NULL: row r — the reference holds one copy per sequence); d_acc2: 2 floats of scratch */
int unimm_t_image_kl(const float* d_logits, int ld, const float* d_target, const int32_t* d_target_row, const int64_t* d_image_label, int rows,
                     int C, float grad_scale, float* d_loss, float* d_dlogits, int ldd, float* d_acc2, void* stream);
/* AdamW of train.py:347 (pytorch_transformers.optimization.AdamW — third-party, not vendored in the reference; restated from its published
 * source): m, v updates, p -= lr sqrt(1 - b2^t) / (1 - b1^t) m / (sqrt(v) + eps) (correct_bias), then p -= lr wd p; g is multiplied by
 * inv_grad_scale first; d_p_lp (optional) receives the refreshed 16-bit operand copy of p. */
int unimm_t_adamw(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, int correct_bias, float inv_grad_scale, void* d_p_lp, int lp_kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIMM_B200_H */
