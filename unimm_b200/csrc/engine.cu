// Engine: owns the packed weights and the activation workspace of one device and runs the layer
// schedule of the reference's BertEncoder.forward (models/vilbert_dialog.py:817-937) followed by the
// heads (:1049-1073), the gathered LM head and the losses (:1559-1624) as a sequence of kernel
// launches on the caller's stream.  C ABI in include/unimm_b200.h.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/unimm_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace unimm {

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;
    int dev = 0;
    UNIMM_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(mu);
    size_t& have = done[{dev, kernel}];
    if (bytes > have) {
        UNIMM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
        have = bytes;
    }
    return 0;
}

namespace {

struct DevTensor {
    float* p = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;
};

struct Linear {
    int N = 0, K = 0;
    const float* w32 = nullptr;  // [N,K] fp32 (fp32 mode, and small heads)
    const bf16* wlp = nullptr;   // [N,K] bf16 (bf16 mode)
    const bf16* wlp_ln = nullptr;  // [N,K] rows permuted for the LayerNorm-fused cluster GEMM (gemm_umma_ln.cu)
    const bf16* wlp_p16 = nullptr; // [N,K] rows permuted for the smem-free 16-bit-output epilogue (gemm_umma.cu)
    int in_kind = LP_BF16;         // encoding of wlp / wlp_ln / wlp_p16, hence of the A operand this projection takes
    const bf16* whl = nullptr;     // [N,2K] fp16 hi | lo planes kept NEXT TO the 16-bit copy (the LM head of the bf16 mode, lm_hp)
    const float* b = nullptr;    // [N]
};
struct LayerNormP {
    const float* g = nullptr;
    const float* b = nullptr;
    int H = 0;
};
struct SelfLayer {  // BertLayer / BertImageLayer
    Linear qkv, out, ffn1, ffn2;
    LayerNormP ln1, ln2;
};
struct ConnLayer {  // BertConnectionLayer
    Linear qkv_v, qkv_t, dense1, dense2, v_ffn1, v_ffn2, t_ffn1, t_ffn2;
    LayerNormP ln1, ln2, v_ln, t_ln;
};

// device-side staging of unimm_score_host (allocated on first use, sized for Bmax)
struct HostPath {
    int64_t *ids = nullptr, *types = nullptr, *pos = nullptr, *labels = nullptr;
    unimm_seq_desc_t* desc = nullptr;
    float *feat = nullptr, *loc = nullptr, *mask = nullptr, *score = nullptr, *nsp = nullptr;
    int32_t *index = nullptr, *rows = nullptr;
    int32_t* h_rows = nullptr;  // pinned
};

// device staging of unimm_score_packed_host: one int32 and one fp32 arena per engine
struct PackedStage {
    int32_t* i32 = nullptr;
    size_t i32_cap = 0;
    float* f32 = nullptr;
    size_t f32_cap = 0;
};

// activation matrix that exists as fp32 and/or bf16
struct ActBuf {
    float* f = nullptr;
    bf16* h = nullptr;
    int ld = 0;
};

}  // namespace
}  // namespace unimm

using namespace unimm;

struct unimm_engine {
    unimm_config_t cfg;
    int device = 0;
    int prec = UNIMM_PREC_FP32;
    int Bmax = 0;
    bool finalized = false;
    std::map<std::string, DevTensor> raw;     // checkpoint tensors, fp32 on device
    std::vector<void*> owned;                 // every cudaMalloc'ed block

    // packed parameters
    const float *word_emb = nullptr, *pos_emb = nullptr, *type_emb = nullptr, *type_ext_emb = nullptr;
    LayerNormP emb_ln, vemb_ln, lm_ln, img_ln;
    Linear img_emb, lm_transform, lm_decoder, img_transform, img_decoder;
    const float *loc_w = nullptr, *loc_b = nullptr;
    const float *nsp_w = nullptr, *nsp_b = nullptr;
    // poolers (ref :946-967): Linear + ReLU on the [CLS] / global-image rows, as fp32-class tcgen05 GEMMs in EVERY precision mode
    // (fp16 hi | lo planes of the fp32 stream and of the weights, three passes: the pooled NSP logit keeps its fp32-level accuracy)
    Linear t_pool, v_pool;
    bf16* pool_split = nullptr;        // [pool_cap, 2 * max(H, Hv)] planes of the pooled rows
    float *pool_t = nullptr, *pool_v = nullptr;   // [pool_cap, Hb]
    int pool_cap = 0;
    int make_pool_linear(const std::string& name, int N, int K, Linear* L);
    // text row of sequence i: t_rows ? t_rows[i] : i * t_stride (same for the image rows)
    int pooler_head(const float* xt_f, int ldt, const int* t_rows, int t_stride, const float* xv_f, int ldv, const int* v_rows, int v_stride,
                    int n, float* nsp, cudaStream_t st);
    int* seq_rows_t = nullptr;         // [Bmax] b * S
    int* seq_rows_v = nullptr;         // [Bmax] b * R
    std::vector<SelfLayer> t_layers, v_layers;
    std::vector<ConnLayer> c_layers;

    // workspace (sized for Bmax sequences)
    ActBuf xt, xv;            // hidden states: fp32 master (+ bf16 shadow in bf16 mode)
    ActBuf xk;                // scores-only packed batches: the labelled rows' stream through the tail of the LAST text layer
    bool xk_live = false;     // the last forward left its final text rows in xk (row i = labelled row i) instead of xt
    float *pre_t = nullptr, *pre_v = nullptr;   // pre-LayerNorm fp32
    void *qkv_t = nullptr, *qkv_v = nullptr, *ctx_t = nullptr, *ctx_v = nullptr, *ffn_t = nullptr, *ffn_v = nullptr;
    void* feat_a = nullptr;        // gathered image features (GEMM operand)
    ActBuf g_in, g_h;                 // LM head rows
    float* g_t1 = nullptr;
    int* g_labels = nullptr;
    float2* partials = nullptr;
    float *label_logit = nullptr, *row_logp = nullptr, *row_ul = nullptr;
    float* lse_u = nullptr;         // log-sum-exp per UNIQUE labelled row (packed batches that share labelled rows)
    bool lm_dedup = true;           // run the LM head once per unique labelled row (UNIMM_LM_DEDUP=0: once per (row, label) entry)
    const int* keep_rows_ = nullptr;   // rows the lean tail of run_encoder keeps (set by forward_packed)
    int n_keep_ = 0;
    float* logits_chunk = nullptr;  // fp32 mode: [kLogitRows, V]
    float* vhead = nullptr;         // image head scratch [Mv, Hv]
    ActBuf vhead_h;
    float* v_logits = nullptr;
    int* err_flag = nullptr;       // device: set by the embedding kernel on an out-of-range token / position / type id
    int* h_err = nullptr;          // pinned copy, read after a stream synchronisation (check_ids)
    HostPath host_path;            // staging of unimm_score_host / unimm_score_packed_host: per engine, so engines on different
    PackedStage packed_stage[2];   // devices can be driven from different threads; two slots: step i + 1 is uploaded and queued
    cudaEvent_t slot_done[2] = {nullptr, nullptr};   // behind step i (unimm_submit_packed_host / unimm_wait_packed)
    cudaStream_t copy_stream = nullptr;              // the upload of step i + 1 runs here, beside step i's kernels
    cudaEvent_t h2d_done[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    int* dense_jobs = nullptr;     // [Bmax, 8] text -> image jobs of the dense layout: (b*S, S, b*R, R, 0, b, 0, 0)
    // host staging for unimm_score_host
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    void* d_stage = nullptr;
    size_t d_stage_bytes = 0;

    static constexpr int kLogitRows = 2048;
    int logits_ld() const { return (cfg.vocab_size + 31) / 32 * 32; }     // 128-byte aligned rows: the coalesced epilogue path

    // optional per-kernel-class timing with CUDA events on the launch stream (bench.py roofline numbers)
    enum { CAT_GEMM = 0, CAT_ATTN = 1, CAT_ROWWISE = 2, CAT_LMHEAD = 3, CAT_GEMM_LN = 4, NCAT = 5 };   // GEMM = umma_gemm_kernel, GEMM_LN = umma_gemm_ln_kernel
    struct ProfRec { cudaEvent_t a, b; int cat; double work; double bytes; };   // bytes = algorithmic HBM bytes of the launch (0: not stated)
    double prof_bytes[5] = {0, 0, 0, 0, 0};    // per class, summed by the last unimm_profile_end
    bool profiling = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    cudaEvent_t prof_event() {
        if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    struct Prof {   // RAII: event before / after whatever is launched in its scope
        unimm_engine* e; cudaStream_t st; ProfRec r; bool on;
        Prof(unimm_engine* e_, int cat, double work, cudaStream_t st_, double bytes = 0.0) : e(e_), st(st_), on(e_->profiling) {
            if (on) { r.a = e->prof_event(); r.b = e->prof_event(); r.cat = cat; r.work = work; r.bytes = bytes; cudaEventRecord(r.a, st); }
        }
        ~Prof() { if (on) { cudaEventRecord(r.b, st); e->prof_recs.push_back(r); } }
    };

    template <typename T>
    int dalloc(T** p, size_t count) {
        void* q = nullptr;
        UNIMM_CUDA_CHECK(cudaMalloc(&q, count * sizeof(T) + 256));
        UNIMM_CUDA_CHECK(cudaMemset(q, 0, count * sizeof(T) + 256));
        owned.push_back(q);
        *p = static_cast<T*>(q);
        return 0;
    }
    bool lp() const { return prec == UNIMM_PREC_BF16 || prec == UNIMM_PREC_FP16; }
    int lp_kind() const { return prec == UNIMM_PREC_FP16 ? LP_FP16 : LP_BF16; }
    size_t esz() const { return lp() ? 2 : 4; }
    // fp32-class mode on the tensor cores: activations and weights as fp16 hi | lo planes (LP_HILO), every projection as three
    // tcgen05 passes into one fp32 accumulator (gemm_umma.cu split3); UNIMM_FP32_SIMT=1 keeps the CUDA-core sgemm instead
    bool tc32_ = false;
    bool tc32() const { return tc32_; }
    // bf16 mode, mixed operand formats (mix16, the default; UNIMM_BF16_PURE=1 turns it off): a LayerNorm output is bounded by
    // |gamma| sqrt(H - 1) + |beta| whatever the checkpoint's activations do, so its 16-bit copy — and the weights of the projections
    // that read it (Q | K | V, FFN-1, the co-attention projections, the LM / image heads) — are fp16: same tcgen05 rate, 3 more
    // mantissa bits.  Everything whose range is NOT bounded by construction (Q / K / V, attention context, GELU outputs, image
    // features, and the weights multiplying them) stays bf16; the residual stream is the fp16 LayerNorm output (res16), fp32 on
    // request.  tcgen05 kind::f16 wants A and B in one format, so the format is a property of the projection (Linear::in_kind),
    // and every GEMM converts on its way out.
    bool mix16 = false;
    int ln_kind() const { return prec == UNIMM_PREC_BF16 && mix16 ? LP_FP16 : lp_kind(); }
    int act_kind() const { return tc32() ? LP_HILO : ln_kind(); }      // what the LayerNorm / embedding kernels write next to fp32
    // UNIMM_LM_HP=1 (16-bit modes with the fp32 residual stream, off by default): the LM head (transform, LayerNorm, vocabulary GEMM +
    // log-sum-exp: 5 % of a scoring step) at the fp32-class precision — the same split3 tcgen05 passes over fp16 hi | lo planes, fed
    // from the fp32 residual stream.  A third of the PURE bf16 mode's error variance is made in this head
    // (tests/bf16_rounding_study.py); with mix16 it buys < 10 % more for 8.5 % of the step, hence a switch and not the default.
    bool lm_hp = false;
    bool lm32() const { return tc32() || (lp() && lm_hp); }
    const bf16* planes_of(const Linear& L) const { return tc32() ? L.wlp : L.whl; }
    bf16* lm_split = nullptr;          // [Mt, 2H] planes of the gathered labelled rows (lm_hp)
    bf16* g_h_hl = nullptr;            // [Mt, 2H] planes of the transformed rows (lm_hp)
    // dense layout in the fp32-class mode: job lists / per-row intervals built from the descriptors (attention_split.cu)
    int *dj_text = nullptr, *dj_i2t = nullptr, *dj_img = nullptr, *dj_row_iv = nullptr;
    // attention over fp16 hi | lo planes: q / k / v point at hi planes of rows `ld` wide whose lo plane lies lo_in columns further;
    // the context goes to o (rows ldo wide, lo plane lo_out further)
    int attention_split(const bf16* q, const bf16* k, const bf16* v, int ld_q, int lo_q, int ld_kv, int lo_kv, bf16* o, int ldo, int lo_out,
                        int heads, int D, const int* jobs, int n_jobs, int max_q, const int* row_iv, const float* key_mask, int key_mask_ld,
                        double qk_pairs, cudaStream_t st);
    bf16* split_scratch = nullptr;     // [rows, 2K] planes of an fp32 operand that no producer wrote as planes
    size_t split_cap = 0;

    int get(const std::string& name, const DevTensor** out, std::vector<int64_t> shape) {
        auto it = raw.find(name);
        UNIMM_CHECK(it != raw.end(), "missing checkpoint key: " + name);
        UNIMM_CHECK(it->second.shape == shape, "unexpected shape for checkpoint key: " + name);
        *out = &it->second;
        return 0;
    }
    // ln_input: the projection's A operand is a LayerNorm output (ln_kind()); otherwise an unbounded 16-bit tensor (lp_kind())
    int make_linear(const std::vector<std::string>& names, int N_each, int K, Linear* L, bool ln_input, bool keep_f32_only = false);
    int make_ln(const std::string& name, int H, LayerNormP* ln);
    int make_ln_weight(Linear* L);
    int finalize();
    int alloc_workspace();

    // y = act(x W^T + b) (+ residual); x/y selected by mode
    int linear(const ActBuf& x, int M, const Linear& L, int act, const float* residual, int ldr, float* out_f32, int ldo_f32,
               void* out_lp, int ldo_lp, cudaStream_t st, int hilo_off = 0);     // hilo_off > 0: fp16 hi | lo planes, lo plane hilo_off columns further
    // out = LayerNorm(x W^T + b + residual): one cluster-fused kernel in the 16-bit modes, GEMM + LayerNorm kernel otherwise
    int linear_ln(const ActBuf& x, int M, const Linear& L, const float* residual, int ldr, const LayerNormP& ln, float* pre,
                  ActBuf& out, cudaStream_t st);
    bool fuse_ln = true;
    bool gelu_tanh = false;      // 1-SFU tanh-form GELU in the FFN-1 epilogue (|err| <= |x| * 2.4e-4): default in fp16 mode, UNIMM_GELU_TANH overrides
    bool prune_tail = true;      // scores-only packed batches: skip everything after the last connection that only the pooled NSP
                                 // logit reads, and run the last text layer's out-proj / FFN on the labelled rows only (UNIMM_PRUNE_TAIL=0 disables)
    bool kv2_ctx_only = true;    // packed layout: co-attention K2 | V2 for the context rows only (UNIMM_KV2_ALL=1 projects every row)
    bool attn_umma = true;       // candidate-row attention on tcgen05 (attention_umma.cu); UNIMM_ATTN_UMMA=0 keeps the mma.sync kernel
    bool frag_epilogue = true;   // QKV / FFN-1 GEMMs read fragment-ordered weight copies (UNIMM_FRAG_EPILOGUE=0 disables)
    // fp16 mode keeps the residual stream in 16 bits between sub-layers (the fused kernel adds it on the tensor core);
    // bf16's 8-bit mantissa cannot afford that, it keeps the fp32 master.  xt.f / xv.f are refreshed after the encoder.
    bool res16 = false;
    int attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B, int heads,
                  int D, int Sq, int Skv, int mask_kind, const SeqDesc* desc, const float* key_mask, cudaStream_t st);
    // which attention a layer runs: dense [B,S]/[B,R] rows with descriptor masks, or jobs over packed rows
    struct AttnCtx {
        const unimm_packed_batch_t* pk = nullptr;   // non-null = packed (prefix-shared) layout
        int B = 0;
        const SeqDesc* desc = nullptr;
        const float* key_mask = nullptr;
        int n_q_rows = 0, n_kv_rows = 0;            // set (per call) to allow the tcgen05 cross-attention kernel
    };
    int attention_packed(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int heads, int D,
                         const int* jobs, int n_jobs, int max_q, int kv_cap, int win_cap, double qk_pairs, const AttnCtx& ac,
                         cudaStream_t st);
    // keep_rows != nullptr: after the attention only these n_keep rows continue (gathered into x_keep, which the layer's output then is)
    int self_layer(const SelfLayer& L, ActBuf& x, float* pre, void* qkv, void* ctx, void* ffn, int M, int heads, bool text,
                   const AttnCtx& ac, cudaStream_t st, const int* keep_rows = nullptr, int n_keep = 0, ActBuf* x_keep = nullptr);
    // image_out = false: the image stream's own update (image->text attention, dense1, image FFN) is skipped — nothing reads it
    int self_layer_tail(const SelfLayer& L, const ActBuf& c, ActBuf& x, float* pre, void* ffn, int M, cudaStream_t st);
    int conn_layer(const ConnLayer& L, int Mt, int Mv, const AttnCtx& ac, cudaStream_t st, bool image_out = true);
    int run_encoder(int Mt, int Mv, const AttnCtx& ac, cudaStream_t st);
    // d_rows == nullptr: src's rows [0, n) are the labelled rows already
    int lm_transform_rows(const ActBuf& src, const int* d_rows, int n, cudaStream_t st);
    int lm_head_rows(const ActBuf& src, const int* d_rows, const int* d_labels, int n, cudaStream_t st);
    // n_u unique rows (of src, or src[d_urows]) serve n (row, label) entries: entry i reads unique row d_uidx[i]
    int lm_head_shared(const ActBuf& src, const int* d_urows, int n_u, const int* d_uidx, const int* d_labels, int n, cudaStream_t st);
    int forward(const unimm_batch_t& in, const unimm_outputs_t& out, cudaStream_t st);
    // the reference's nn.Embedding raises on an id outside its table (models/vilbert_dialog.py:334-350); the embedding kernel clamps
    // and raises err_flag instead: queue its copy behind the forward, and after the caller's synchronisation turn it into an error
    int queue_id_check(cudaStream_t st, int slot = 0) {
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(h_err + slot, err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        UNIMM_CUDA_CHECK(cudaMemsetAsync(err_flag, 0, sizeof(int), st));      // sticky between checks, cleared by each one
        return 0;
    }
    int id_check_result(int slot = 0) {
        const int bad = h_err[slot];
        h_err[slot] = 0;
        UNIMM_CHECK(bad == 0, "token, position or token-type id outside its embedding table (the reference raises an IndexError here)");
        return 0;
    }
    int forward_packed(const unimm_packed_batch_t& in, float* d_seq_score, float* d_nsp_scores, float* d_token_logp, cudaStream_t st);
};

namespace {
inline char* byte_ptr(void* p) { return static_cast<char*>(p); }
}  // namespace

int unimm_engine::make_ln(const std::string& name, int H, LayerNormP* ln) {
    const DevTensor *g, *b;
    UNIMM_TRY(get(name + ".weight", &g, {H}));
    UNIMM_TRY(get(name + ".bias", &b, {H}));
    ln->g = g->p; ln->b = b->p; ln->H = H;
    return 0;
}

// Concatenate one or more nn.Linear weights along the output dimension ([sum N, K]) and, in bf16 mode,
// cast the packed matrix once.
int unimm_engine::make_linear(const std::vector<std::string>& names, int N_each, int K, Linear* L, bool ln_input, bool keep_f32_only) {
    const int parts = static_cast<int>(names.size());
    L->in_kind = ln_input ? ln_kind() : lp_kind();
    L->N = N_each * parts;
    L->K = K;
    float *w = nullptr, *b = nullptr;
    if (parts == 1) {
        const DevTensor *tw, *tb;
        UNIMM_TRY(get(names[0] + ".weight", &tw, {N_each, K}));
        UNIMM_TRY(get(names[0] + ".bias", &tb, {N_each}));
        w = tw->p; b = tb->p;
    } else {
        UNIMM_TRY(dalloc(&w, static_cast<size_t>(L->N) * K));
        UNIMM_TRY(dalloc(&b, static_cast<size_t>(L->N)));
        for (int i = 0; i < parts; ++i) {
            const DevTensor *tw, *tb;
            UNIMM_TRY(get(names[i] + ".weight", &tw, {N_each, K}));
            UNIMM_TRY(get(names[i] + ".bias", &tb, {N_each}));
            UNIMM_CUDA_CHECK(cudaMemcpy(w + static_cast<size_t>(i) * N_each * K, tw->p, sizeof(float) * N_each * K, cudaMemcpyDeviceToDevice));
            UNIMM_CUDA_CHECK(cudaMemcpy(b + static_cast<size_t>(i) * N_each, tb->p, sizeof(float) * N_each, cudaMemcpyDeviceToDevice));
        }
    }
    L->w32 = w;
    L->b = b;
    if (tc32() && !keep_f32_only && K % 64 == 0) {
        bf16* h = nullptr;
        UNIMM_TRY(dalloc(&h, static_cast<size_t>(L->N) * 2 * K));
        UNIMM_TRY(split_f32_to_hilo(w, K, L->N, K, h, 0));
        L->wlp = h;
    }
    if (lp() && !keep_f32_only) {
        bf16* h = nullptr;
        UNIMM_TRY(dalloc(&h, static_cast<size_t>(L->N) * K));
        UNIMM_TRY(cast_f32_to_lp(w, h, static_cast<size_t>(L->N) * K, L->in_kind, 0));
        L->wlp = h;
        if (frag_epilogue && L->N % 32 == 0 && K % 8 == 0) {
            bf16* hp = nullptr;
            UNIMM_TRY(dalloc(&hp, static_cast<size_t>(L->N) * K));
            UNIMM_TRY(permute_weight_rows(h, hp, L->N, K, 1, 0));
            L->wlp_p16 = hp;
        }
    }
    return 0;
}

int unimm_engine::make_pool_linear(const std::string& name, int N, int K, Linear* L) {
    const DevTensor *tw, *tb;
    UNIMM_TRY(get(name + ".weight", &tw, {N, K}));
    UNIMM_TRY(get(name + ".bias", &tb, {N}));
    UNIMM_CHECK(K % 64 == 0, "pooler input width must be a multiple of 64");
    L->N = N; L->K = K; L->w32 = tw->p; L->b = tb->p;
    bf16* h = nullptr;
    UNIMM_TRY(dalloc(&h, static_cast<size_t>(N) * 2 * K));
    UNIMM_TRY(split_f32_to_hilo(tw->p, K, N, K, h, 0));
    L->wlp = h;
    return 0;
}

// poolers + NSP head: pooled = relu(W x_row + b) for the text and the image row of each sequence as two split3 GEMMs (M = n),
// then nsp = Wn (pooled_t * pooled_v) + bn
int unimm_engine::pooler_head(const float* xt_f, int ldt, const int* t_rows, int t_stride, const float* xv_f, int ldv, const int* v_rows,
                              int v_stride, int n, float* nsp, cudaStream_t st) {
    const int Hb = cfg.bi_hidden_size;
    UNIMM_CHECK(t_rows != nullptr || t_stride == cfg.seq_len, "pooler: dense rows are b * S");
    for (int s0 = 0; s0 < n; s0 += pool_cap) {
        const int m = std::min(pool_cap, n - s0);
        for (int side = 0; side < 2; ++side) {
            const Linear& L = side == 0 ? t_pool : v_pool;
            const float* x = side == 0 ? xt_f : xv_f;
            const int ld = side == 0 ? ldt : ldv;
            const int* rows = side == 0 ? t_rows : v_rows;
            const int* rows_dense = side == 0 ? seq_rows_t : seq_rows_v;
            UNIMM_TRY(split_f32_to_hilo(x, ld, m, L.K, pool_split, st, rows != nullptr ? rows + s0 : rows_dense + s0));
            GemmEpilogue ep;
            ep.bias = L.b; ep.act = ACT_RELU; ep.lp_kind = LP_FP16; ep.split3 = 1;
            ep.out_f32 = side == 0 ? pool_t : pool_v; ep.ldo_f32 = Hb;
            UNIMM_TRY(gemm_umma_bf16(pool_split, 2 * L.K, L.wlp, 2 * L.K, m, Hb, L.K, ep, 0, 0, st));
        }
        UNIMM_TRY(nsp_from_pooled(pool_t, pool_v, m, Hb, nsp_w, nsp_b, nsp + static_cast<size_t>(s0) * 2, st));
    }
    return 0;
}

// second 16-bit copy of a weight whose output feeds a residual + LayerNorm, in the row order the fused kernel wants
int unimm_engine::make_ln_weight(Linear* L) {
    if (!lp() || !fuse_ln || L->wlp == nullptr || (L->N != 768 && L->N != 1024) || L->K % 64 != 0) return 0;
    bf16* h = nullptr;
    UNIMM_TRY(dalloc(&h, static_cast<size_t>(L->N) * L->K));
    UNIMM_TRY(permute_weight_rows_ln(L->wlp, h, L->N, L->K, 0));
    L->wlp_ln = h;
    return 0;
}

int unimm_engine::finalize() {
    UNIMM_CHECK(!finalized, "weights already finalized");
    const unimm_config_t& c = cfg;
    const int H = c.hidden_size, Hv = c.v_hidden_size, Hb = c.bi_hidden_size, I = c.intermediate_size, Iv = c.v_intermediate_size;
    const DevTensor* t;
    UNIMM_TRY(get("bert.embeddings.word_embeddings.weight", &t, {c.vocab_size, H})); word_emb = t->p;
    UNIMM_TRY(get("bert.embeddings.position_embeddings.weight", &t, {c.max_position_embeddings, H})); pos_emb = t->p;
    UNIMM_TRY(get("bert.embeddings.token_type_embeddings.weight", &t, {c.type_vocab_size, H})); type_emb = t->p;
    UNIMM_TRY(get("bert.embeddings.token_type_embeddings_extension.weight", &t, {10, H})); type_ext_emb = t->p;
    UNIMM_TRY(make_ln("bert.embeddings.LayerNorm", H, &emb_ln));
    UNIMM_TRY(make_linear({"bert.v_embeddings.image_embeddings"}, Hv, c.v_feature_size, &img_emb, false));
    UNIMM_TRY(get("bert.v_embeddings.image_location_embeddings.weight", &t, {Hv, 5})); loc_w = t->p;
    UNIMM_TRY(get("bert.v_embeddings.image_location_embeddings.bias", &t, {Hv})); loc_b = t->p;
    UNIMM_TRY(make_ln("bert.v_embeddings.LayerNorm", Hv, &vemb_ln));

    t_layers.resize(c.num_hidden_layers);
    for (int i = 0; i < c.num_hidden_layers; ++i) {
        const std::string p = "bert.encoder.layer." + std::to_string(i) + ".";
        SelfLayer& L = t_layers[i];
        UNIMM_TRY(make_linear({p + "attention.self.query", p + "attention.self.key", p + "attention.self.value"}, H, H, &L.qkv, true));
        UNIMM_TRY(make_linear({p + "attention.output.dense"}, H, H, &L.out, false));
        UNIMM_TRY(make_ln(p + "attention.output.LayerNorm", H, &L.ln1));
        UNIMM_TRY(make_linear({p + "intermediate.dense"}, I, H, &L.ffn1, true));
        UNIMM_TRY(make_linear({p + "output.dense"}, H, I, &L.ffn2, false));
        UNIMM_TRY(make_ln(p + "output.LayerNorm", H, &L.ln2));
    }
    v_layers.resize(c.v_num_hidden_layers);
    for (int i = 0; i < c.v_num_hidden_layers; ++i) {
        const std::string p = "bert.encoder.v_layer." + std::to_string(i) + ".";
        SelfLayer& L = v_layers[i];
        UNIMM_TRY(make_linear({p + "attention.self.query", p + "attention.self.key", p + "attention.self.value"}, Hv, Hv, &L.qkv, true));
        UNIMM_TRY(make_linear({p + "attention.output.dense"}, Hv, Hv, &L.out, false));
        UNIMM_TRY(make_ln(p + "attention.output.LayerNorm", Hv, &L.ln1));
        UNIMM_TRY(make_linear({p + "intermediate.dense"}, Iv, Hv, &L.ffn1, true));
        UNIMM_TRY(make_linear({p + "output.dense"}, Hv, Iv, &L.ffn2, false));
        UNIMM_TRY(make_ln(p + "output.LayerNorm", Hv, &L.ln2));
    }
    c_layers.resize(c.num_connections);
    for (int i = 0; i < c.num_connections; ++i) {
        const std::string p = "bert.encoder.c_layer." + std::to_string(i) + ".";
        ConnLayer& L = c_layers[i];
        UNIMM_TRY(make_linear({p + "biattention.query1", p + "biattention.key1", p + "biattention.value1"}, Hb, Hv, &L.qkv_v, true));
        UNIMM_TRY(make_linear({p + "biattention.query2", p + "biattention.key2", p + "biattention.value2"}, Hb, H, &L.qkv_t, true));
        UNIMM_TRY(make_linear({p + "biOutput.dense1"}, Hv, Hb, &L.dense1, false));
        UNIMM_TRY(make_ln(p + "biOutput.LayerNorm1", Hv, &L.ln1));
        UNIMM_TRY(make_linear({p + "biOutput.dense2"}, H, Hb, &L.dense2, false));
        UNIMM_TRY(make_ln(p + "biOutput.LayerNorm2", H, &L.ln2));
        UNIMM_TRY(make_linear({p + "v_intermediate.dense"}, Iv, Hv, &L.v_ffn1, true));
        UNIMM_TRY(make_linear({p + "v_output.dense"}, Hv, Iv, &L.v_ffn2, false));
        UNIMM_TRY(make_ln(p + "v_output.LayerNorm", Hv, &L.v_ln));
        UNIMM_TRY(make_linear({p + "t_intermediate.dense"}, I, H, &L.t_ffn1, true));
        UNIMM_TRY(make_linear({p + "t_output.dense"}, H, I, &L.t_ffn2, false));
        UNIMM_TRY(make_ln(p + "t_output.LayerNorm", H, &L.t_ln));
    }
    UNIMM_TRY(make_pool_linear("bert.t_pooler.dense", Hb, H, &t_pool));
    UNIMM_TRY(make_pool_linear("bert.v_pooler.dense", Hb, Hv, &v_pool));
    UNIMM_TRY(get("cls.bi_seq_relationship.weight", &t, {2, Hb})); nsp_w = t->p;
    UNIMM_TRY(get("cls.bi_seq_relationship.bias", &t, {2})); nsp_b = t->p;

    UNIMM_TRY(make_linear({"cls.predictions.transform.dense"}, H, H, &lm_transform, true));
    if (lp() && lm_hp) {
        UNIMM_CHECK(H % 64 == 0 && !(fuse_ln && res16), "the fp32-class LM head needs H % 64 == 0 and the fp32 residual stream");
        bf16* h = nullptr;
        UNIMM_TRY(dalloc(&h, static_cast<size_t>(H) * 2 * H));
        UNIMM_TRY(split_f32_to_hilo(lm_transform.w32, H, H, H, h, 0));
        lm_transform.whl = h;
    }
    UNIMM_TRY(make_ln("cls.predictions.transform.LayerNorm", H, &lm_ln));
    // tied decoder (reference :1020): accept either key, insist they agree when both are present (compared below)
    {
        lm_decoder.N = c.vocab_size;
        lm_decoder.K = H;
        lm_decoder.w32 = word_emb;
        auto it = raw.find("cls.predictions.decoder.weight");
        if (it != raw.end()) {
            UNIMM_CHECK(it->second.numel == static_cast<size_t>(c.vocab_size) * H, "bad decoder weight shape");
            // the reference ties the two (one Parameter, :1020): a checkpoint in which they differ is not a reference checkpoint
            int* d_diff = nullptr;
            UNIMM_TRY(dalloc(&d_diff, 4));
            UNIMM_TRY(buffers_differ(it->second.p, word_emb, it->second.numel, d_diff, 0));
            int h_diff = 0;
            UNIMM_CUDA_CHECK(cudaMemcpy(&h_diff, d_diff, sizeof(int), cudaMemcpyDeviceToHost));
            UNIMM_CHECK(h_diff == 0, "cls.predictions.decoder.weight differs from bert.embeddings.word_embeddings.weight (the reference ties them)");
        }
        UNIMM_TRY(get("cls.predictions.bias", &t, {c.vocab_size}));
        lm_decoder.b = t->p;
        if (lp()) {
            bf16* h = nullptr;
            UNIMM_TRY(dalloc(&h, static_cast<size_t>(c.vocab_size) * H));
            lm_decoder.in_kind = ln_kind();
            UNIMM_TRY(cast_f32_to_lp(lm_decoder.w32, h, static_cast<size_t>(c.vocab_size) * H, ln_kind(), 0));
            lm_decoder.wlp = h;
            if (lm_hp) {
                bf16* hl = nullptr;
                UNIMM_TRY(dalloc(&hl, static_cast<size_t>(c.vocab_size) * 2 * H));
                UNIMM_TRY(split_f32_to_hilo(lm_decoder.w32, H, c.vocab_size, H, hl, 0));
                lm_decoder.whl = hl;
            }
        } else if (tc32()) {
            bf16* h = nullptr;
            UNIMM_TRY(dalloc(&h, static_cast<size_t>(c.vocab_size) * 2 * H));
            UNIMM_TRY(split_f32_to_hilo(lm_decoder.w32, H, c.vocab_size, H, h, 0));
            lm_decoder.wlp = h;
        }
    }
    UNIMM_TRY(make_linear({"cls.imagePredictions.transform.dense"}, Hv, Hv, &img_transform, true));
    UNIMM_TRY(make_ln("cls.imagePredictions.transform.LayerNorm", Hv, &img_ln));
    UNIMM_TRY(make_linear({"cls.imagePredictions.decoder"}, c.v_target_size, Hv, &img_decoder, true));
    UNIMM_TRY(make_ln_weight(&img_emb));
    for (auto& L : t_layers) { UNIMM_TRY(make_ln_weight(&L.out)); UNIMM_TRY(make_ln_weight(&L.ffn2)); }
    for (auto& L : v_layers) { UNIMM_TRY(make_ln_weight(&L.out)); UNIMM_TRY(make_ln_weight(&L.ffn2)); }
    for (auto& L : c_layers) {
        UNIMM_TRY(make_ln_weight(&L.dense1)); UNIMM_TRY(make_ln_weight(&L.dense2));
        UNIMM_TRY(make_ln_weight(&L.v_ffn2)); UNIMM_TRY(make_ln_weight(&L.t_ffn2));
    }
    UNIMM_CUDA_CHECK(cudaDeviceSynchronize());
    UNIMM_TRY(alloc_workspace());
    finalized = true;
    return 0;
}

int unimm_engine::alloc_workspace() {
    const unimm_config_t& c = cfg;
    const size_t Mt = static_cast<size_t>(Bmax) * c.seq_len, Mv = static_cast<size_t>(Bmax) * c.num_regions;
    const int H = c.hidden_size, Hv = c.v_hidden_size, Hb = c.bi_hidden_size;
    const size_t wt = std::max(std::max(3 * H, 3 * Hb), c.intermediate_size);  // widest text-row buffer
    const size_t wv = std::max(std::max(3 * Hv, 3 * Hb), c.v_intermediate_size);
    UNIMM_TRY(dalloc(&xt.f, Mt * H)); xt.ld = H;
    UNIMM_TRY(dalloc(&xv.f, Mv * Hv)); xv.ld = Hv;
    const size_t planes = tc32() ? 2 : 1;           // LP_HILO rows are two fp16 planes wide
    if (lp() || tc32()) { UNIMM_TRY(dalloc(&xt.h, Mt * H * planes)); UNIMM_TRY(dalloc(&xv.h, Mv * Hv * planes)); }
    UNIMM_TRY(dalloc(&xk.f, Mt * H)); xk.ld = H;
    if (lp() || tc32()) UNIMM_TRY(dalloc(&xk.h, Mt * H * planes));
    if (tc32()) {
        UNIMM_TRY(dalloc(&dj_text, static_cast<size_t>(Bmax) * 8));
        UNIMM_TRY(dalloc(&dj_i2t, static_cast<size_t>(Bmax) * 8));
        UNIMM_TRY(dalloc(&dj_img, static_cast<size_t>(Bmax) * 8));
        UNIMM_TRY(dalloc(&dj_row_iv, Mt * 4));
        split_cap = std::max(Mt * 2 * static_cast<size_t>(std::max(std::max(H, Hb), Hv)), Mv * 2 * static_cast<size_t>(std::max(c.v_feature_size, std::max(Hv, Hb))));
        UNIMM_TRY(dalloc(&split_scratch, split_cap));
    }
    UNIMM_TRY(dalloc(&pre_t, Mt * H));
    UNIMM_TRY(dalloc(&pre_v, Mv * Hv));
    char* p;
    UNIMM_TRY(dalloc(&p, Mt * wt * esz())); qkv_t = p;
    UNIMM_TRY(dalloc(&p, Mv * wv * esz())); qkv_v = p;
    UNIMM_TRY(dalloc(&p, Mt * std::max(H, Hb) * esz())); ctx_t = p;
    UNIMM_TRY(dalloc(&p, Mv * std::max(Hv, Hb) * esz())); ctx_v = p;
    UNIMM_TRY(dalloc(&p, Mt * wt * esz())); ffn_t = p;
    UNIMM_TRY(dalloc(&p, Mv * wv * esz())); ffn_v = p;
    UNIMM_TRY(dalloc(&p, Mv * c.v_feature_size * esz())); feat_a = p;
    // LM head rows: worst case every text position is labelled
    UNIMM_TRY(dalloc(&g_in.f, lp() ? 4 : Mt * H)); g_in.ld = H;
    UNIMM_TRY(dalloc(&g_h.f, Mt * H)); g_h.ld = H;
    if (lp()) { UNIMM_TRY(dalloc(&g_in.h, Mt * H)); UNIMM_TRY(dalloc(&g_h.h, Mt * H)); }
    if (tc32()) UNIMM_TRY(dalloc(&g_h.h, Mt * H * 2));
    if (lp() && lm_hp) { UNIMM_TRY(dalloc(&lm_split, Mt * H * 2)); UNIMM_TRY(dalloc(&g_h_hl, Mt * H * 2)); }
    UNIMM_TRY(dalloc(&g_t1, Mt * H));
    UNIMM_TRY(dalloc(&g_labels, Mt));
    UNIMM_TRY(dalloc(&label_logit, Mt));
    UNIMM_TRY(dalloc(&row_logp, Mt));
    UNIMM_TRY(dalloc(&row_ul, Mt));
    UNIMM_TRY(dalloc(&lse_u, Mt));
    if (lp() || tc32()) UNIMM_TRY(dalloc(&partials, Mt * gemm_umma_lse_tiles(c.vocab_size)));
    if (!lp() && !tc32()) UNIMM_TRY(dalloc(&logits_chunk, static_cast<size_t>(kLogitRows) * logits_ld()));
    UNIMM_TRY(dalloc(&vhead, Mv * Hv));
    vhead_h.ld = Hv;
    UNIMM_TRY(dalloc(&vhead_h.f, Mv * Hv));
    if (lp() || tc32()) UNIMM_TRY(dalloc(&vhead_h.h, Mv * Hv * planes));
    UNIMM_TRY(dalloc(&v_logits, Mv * c.v_target_size));
    {
        pool_cap = std::max(Bmax, 2048);
        UNIMM_TRY(dalloc(&pool_split, static_cast<size_t>(pool_cap) * 2 * std::max(H, Hv)));
        UNIMM_TRY(dalloc(&pool_t, static_cast<size_t>(pool_cap) * Hb));
        UNIMM_TRY(dalloc(&pool_v, static_cast<size_t>(pool_cap) * Hb));
        std::vector<int> rt(pool_cap), rv(pool_cap);
        for (int b = 0; b < pool_cap; ++b) { rt[b] = b * c.seq_len; rv[b] = b * c.num_regions; }
        UNIMM_TRY(dalloc(&seq_rows_t, pool_cap));
        UNIMM_TRY(dalloc(&seq_rows_v, pool_cap));
        UNIMM_CUDA_CHECK(cudaMemcpy(seq_rows_t, rt.data(), sizeof(int) * pool_cap, cudaMemcpyHostToDevice));
        UNIMM_CUDA_CHECK(cudaMemcpy(seq_rows_v, rv.data(), sizeof(int) * pool_cap, cudaMemcpyHostToDevice));
    }
    UNIMM_TRY(dalloc(&err_flag, 4));
    UNIMM_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&h_err), 4 * sizeof(int)));
    h_err[0] = h_err[1] = h_err[2] = h_err[3] = 0;
    {
        std::vector<int> jobs(static_cast<size_t>(Bmax) * 8, 0);
        for (int b = 0; b < Bmax; ++b) {
            int* j = &jobs[static_cast<size_t>(b) * 8];
            j[0] = b * c.seq_len; j[1] = c.seq_len; j[2] = b * c.num_regions; j[3] = c.num_regions; j[4] = 0; j[5] = b;
        }
        UNIMM_TRY(dalloc(&dense_jobs, jobs.size()));
        UNIMM_CUDA_CHECK(cudaMemcpy(dense_jobs, jobs.data(), jobs.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return 0;
}

int unimm_engine::linear(const ActBuf& x, int M, const Linear& L, int act, const float* residual, int ldr, float* out_f32,
                         int ldo_f32, void* out_lp, int ldo_lp, cudaStream_t st, int hilo_off) {
    const bool out_hilo = hilo_off > 0;
    GemmEpilogue ep;
    ep.bias = L.b;
    ep.residual = residual;
    ep.ldr = ldr;
    ep.act = act;
    ep.lp_kind = lp_kind();
    if (lp()) ep.in_kind = L.in_kind;
    // algorithmic bytes: A + W once, every output (and the residual) once
    const double mn = static_cast<double>(M) * L.N;
    Prof prof(this, CAT_GEMM, 2.0 * M * L.N * L.K, st,
              esz() * (static_cast<double>(M) * L.K + static_cast<double>(L.N) * L.K) + (out_lp ? esz() * mn : 0.0) +
                  (out_f32 ? 4.0 * mn : 0.0) + (residual ? 4.0 * mn : 0.0));
    if (lp()) {
        ep.out_f32 = out_f32; ep.ldo_f32 = ldo_f32;
        ep.out_bf16 = static_cast<bf16*>(out_lp); ep.ldo_bf16 = ldo_lp;
        UNIMM_CHECK(x.h != nullptr && L.wlp != nullptr, "bf16 operand missing");
        if (L.wlp_p16 != nullptr && out_f32 == nullptr && residual == nullptr && out_lp != nullptr) {
            ep.w_perm16 = true;
            if (act == ACT_GELU && gelu_tanh) ep.act = ACT_GELU_TANH;
            return gemm_umma_bf16(x.h, x.ld, L.wlp_p16, L.K, M, L.N, L.K, ep, 0, 0, st);
        }
        return gemm_umma_bf16(x.h, x.ld, L.wlp, L.K, M, L.N, L.K, ep, 0, 0, st);
    }
    // fp32 mode: the "low precision" output slot is an fp32 tensor as well (or, out_hilo, its fp16 hi | lo planes)
    UNIMM_CHECK(!(out_f32 && out_lp), "fp32 mode writes one output");
    if (out_hilo) {
        UNIMM_CHECK(tc32() && out_lp != nullptr, "hi | lo output is the fp32-class tensor-core mode's");
        ep.out_bf16 = static_cast<bf16*>(out_lp); ep.ldo_bf16 = 2 * ldo_lp; ep.out_hilo = true; ep.hilo_off = hilo_off;
    } else {
        ep.out_f32 = out_f32 ? out_f32 : static_cast<float*>(out_lp);
        ep.ldo_f32 = out_f32 ? ldo_f32 : ldo_lp;
    }
    if (tc32() && L.wlp != nullptr && L.K % 64 == 0) {
        const bf16* a = x.h;
        if (a == nullptr) {            // an fp32 operand nobody wrote as planes (attention context, gathered rows, image features)
            UNIMM_CHECK(x.f != nullptr && static_cast<size_t>(M) * 2 * L.K <= split_cap, "fp32-class GEMM: operand larger than the split scratch");
            Prof prof2(this, CAT_ROWWISE, 8.0 * M * L.K, st);
            UNIMM_TRY(split_f32_to_hilo(x.f, x.ld, M, L.K, split_scratch, st));
            a = split_scratch;
        } else {
            UNIMM_CHECK(x.ld == L.K, "fp32-class GEMM: plane operand with a leading dimension other than K");
        }
        ep.lp_kind = LP_FP16;
        ep.split3 = 1;
        if (act == ACT_GELU) ep.act = ACT_GELU_ERF;
        return gemm_umma_bf16(a, 2 * L.K, L.wlp, 2 * L.K, M, L.N, L.K, ep, 0, 0, st);
    }
    UNIMM_CHECK(!out_hilo, "hi | lo output needs the tensor-core path (K % 64 == 0)");
    return gemm_simt_f32(x.f, x.ld, L.w32, L.K, M, L.N, L.K, ep, st);
}

int unimm_engine::linear_ln(const ActBuf& x, int M, const Linear& L, const float* residual, int ldr, const LayerNormP& ln, float* pre,
                            ActBuf& out, cudaStream_t st) {
    if (lp() && fuse_ln) {
        GemmLnEpilogue ep;
        ep.bias = L.b; ep.residual = residual; ep.ldr = ldr; ep.gamma = ln.g; ep.beta = ln.b;
        ep.out_f32 = out.f; ep.ldo_f32 = out.ld; ep.out_lp = out.h; ep.ldo_lp = out.ld; ep.lp_kind = ln_kind(); ep.in_kind = L.in_kind;
        if (res16 && residual == out.f && ldr == out.ld) {   // the stream's own previous value: use (and update) its 16-bit copy only
            ep.residual = nullptr; ep.residual_lp = out.h; ep.ldr_lp = out.ld; ep.out_f32 = nullptr;
        }
        if (L.wlp_ln != nullptr && gemm_umma_ln_supported(L.N, L.K, ep)) {
            UNIMM_CHECK(x.h != nullptr, "16-bit operand missing");
            Prof prof(this, CAT_GEMM_LN, 2.0 * M * L.N * L.K, st);
            return gemm_umma_ln(x.h, x.ld, L.wlp_ln, L.K, M, L.N, L.K, ep, st);
        }
    }
    UNIMM_CHECK(!(lp() && fuse_ln && res16), "16-bit residual stream needs the LayerNorm-fused GEMM (N = 768 or 1024, K % 64 == 0)");
    UNIMM_TRY(linear(x, M, L, ACT_NONE, residual, ldr, pre, L.N, nullptr, 0, st));
    Prof prof(this, CAT_ROWWISE, (8.0 + esz()) * M * L.N, st);
    return layernorm_rows(pre, L.N, M, L.N, ln.g, ln.b, out.f, out.h, act_kind(), st);
}

int unimm_engine::attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B,
                            int heads, int D, int Sq, int Skv, int mask_kind, const SeqDesc* desc, const float* key_mask,
                            cudaStream_t st) {
    AttnArgs a;
    a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.o = o; a.ldo = ldo;
    a.B = B; a.heads = heads; a.D = D; a.Sq = Sq; a.Skv = Skv;
    a.mask_kind = mask_kind; a.desc = desc; a.key_mask = key_mask;
    a.scale = 1.0f / sqrtf(static_cast<float>(D));
    a.lp_kind = lp_kind();
    Prof prof(this, CAT_ATTN, 4.0 * B * heads * static_cast<double>(Sq) * Skv * D, st);
    if (lp() && attn_umma) {
        AttnJobsArgs j;
        j.q = q; j.ldq = ldq; j.k = k; j.ldk = ldk; j.v = v; j.ldv = ldv; j.o = o; j.ldo = ldo;
        j.heads = heads; j.D = D; j.jobs = nullptr; j.n_jobs = B; j.max_q_len = Sq; j.kv_cap = Skv <= 64 ? 64 : 256; j.win_cap = 0;
        j.row_iv = nullptr; j.key_mask = nullptr; j.key_mask_ld = 0; j.scale = a.scale; j.lp_kind = lp_kind();
        if (mask_kind == MASK_TEXT_SELF && Sq == Skv) {          // text self-attention of the dense layout: tcgen05, masks from descriptors
            j.n_rows = B * Sq; j.desc = desc; j.seq_len = Sq;
            if (attention_dense_umma_supported(j)) return attention_dense_umma(j, st);
        } else if (mask_kind == MASK_KEY_VECTOR && Sq == cfg.seq_len && Skv == cfg.num_regions && Sq > 64 && key_mask != nullptr) {
            // text -> image: one job per sequence (rows b*S.., keys b*R.., mask row b)
            j.jobs = dense_jobs; j.n_rows = B * Sq; j.n_kv_rows = B * Skv; j.key_mask = key_mask; j.key_mask_ld = Skv;
            if (attention_cross_umma_supported(j)) return attention_cross_umma(j, st);
        }
    }
    return lp() ? attention_mma_lp(a, st) : attention_simt_f32(a, st);
}

int unimm_engine::attention_split(const bf16* q, const bf16* k, const bf16* v, int ld_q, int lo_q, int ld_kv, int lo_kv, bf16* o, int ldo,
                                  int lo_out, int heads, int D, const int* jobs, int n_jobs, int max_q, const int* row_iv,
                                  const float* key_mask, int key_mask_ld, double qk_pairs, cudaStream_t st) {
    AttnJobsArgs a;
    a.q = q; a.ldq = ld_q; a.k = k; a.ldk = ld_kv; a.v = v; a.ldv = ld_kv; a.o = o; a.ldo = ldo;
    a.lo_off_q = lo_q; a.lo_off_k = lo_kv; a.lo_off_v = lo_kv; a.lo_off_o = lo_out;
    a.heads = heads; a.D = D; a.jobs = jobs; a.n_jobs = n_jobs; a.max_q_len = max_q; a.kv_cap = 256; a.win_cap = 256;
    a.row_iv = row_iv; a.key_mask = key_mask; a.key_mask_ld = key_mask_ld;
    a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = LP_FP16;
    Prof prof(this, CAT_ATTN, 4.0 * heads * D * qk_pairs, st);
    return attention_jobs_split(a, st);
}

int unimm_engine::attention_packed(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int heads,
                                   int D, const int* jobs, int n_jobs, int max_q, int kv_cap, int win_cap, double qk_pairs,
                                   const AttnCtx& ac, cudaStream_t st) {
    AttnJobsArgs a;
    a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.o = o; a.ldo = ldo;
    a.heads = heads; a.D = D; a.jobs = jobs; a.n_jobs = n_jobs; a.max_q_len = max_q; a.kv_cap = kv_cap; a.win_cap = win_cap;
    a.row_iv = ac.pk->d_row_iv; a.key_mask = ac.pk->d_image_mask; a.key_mask_ld = cfg.num_regions;
    a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind();
    a.n_rows = ac.n_q_rows; a.n_kv_rows = ac.n_kv_rows;
    Prof prof(this, CAT_ATTN, 4.0 * heads * D * qk_pairs, st);
    if (lp() && attn_umma && ac.n_q_rows > 0 && max_q > 64 && attention_cross_umma_supported(a)) return attention_cross_umma(a, st);
    return attention_jobs(a, !lp(), st);
}

// BertLayer / BertImageLayer: QKV -> attention -> out-proj + residual -> LN -> FFN1+GELU -> FFN2 + residual -> LN
int unimm_engine::self_layer(const SelfLayer& L, ActBuf& x_in, float* pre, void* qkv, void* ctx, void* ffn, int M_in, int heads, bool text,
                             const AttnCtx& ac, cudaStream_t st, const int* keep_rows, int n_keep, ActBuf* x_keep) {
    ActBuf& x = x_in;
    const int M = M_in;
    const int H = x.ld, D = H / heads;
    const size_t e = esz();
    if (tc32()) {
        // fp32-class mode: Q | K | V as fp16 hi | lo planes ([M, 6H]: hi planes in columns [0, 3H), lo planes 3H further) straight
        // from the split3 GEMM, attention as three mma.sync passes per product, context out as planes [M, 2H]
        UNIMM_TRY(linear(x, M, L.qkv, ACT_NONE, nullptr, 0, nullptr, 0, qkv, 3 * H, st, 3 * H));
        const bf16* q16 = static_cast<const bf16*>(qkv);
        bf16* c16 = static_cast<bf16*>(ctx);
        const int S = cfg.seq_len, R = cfg.num_regions;
        if (ac.pk != nullptr) {
            const unimm_packed_batch_t& pk = *ac.pk;
            if (text) UNIMM_TRY(attention_split(q16, q16 + H, q16 + 2 * H, 6 * H, 3 * H, 6 * H, 3 * H, c16, 2 * H, H, heads, D, pk.d_jobs_text_self,
                                                pk.n_jobs_text_self, pk.max_q_text_self, pk.d_row_iv, nullptr, 0, pk.pairs_text_self, st));
            else UNIMM_TRY(attention_split(q16, q16 + H, q16 + 2 * H, 6 * H, 3 * H, 6 * H, 3 * H, c16, 2 * H, H, heads, D, pk.d_jobs_img_self,
                                           pk.n_jobs_img_self, R, nullptr, pk.d_image_mask, R, static_cast<double>(pk.n_units) * R * R, st));
        } else {
            if (text) UNIMM_TRY(attention_split(q16, q16 + H, q16 + 2 * H, 6 * H, 3 * H, 6 * H, 3 * H, c16, 2 * H, H, heads, D, dj_text, ac.B, S,
                                                dj_row_iv, nullptr, 0, static_cast<double>(ac.B) * S * S, st));
            else UNIMM_TRY(attention_split(q16, q16 + H, q16 + 2 * H, 6 * H, 3 * H, 6 * H, 3 * H, c16, 2 * H, H, heads, D, dj_img, ac.B, R, nullptr,
                                           ac.key_mask, R, static_cast<double>(ac.B) * R * R, st));
        }
        ActBuf c;
        c.f = nullptr; c.h = c16; c.ld = H;
        if (keep_rows != nullptr) {
            UNIMM_CHECK(x_keep != nullptr && n_keep > 0 && n_keep <= M, "self_layer: bad row subset");
            ActBuf cc;
            cc.f = nullptr; cc.h = static_cast<bf16*>(ffn); cc.ld = H;
            {
                Prof prof(this, CAT_ROWWISE, 16.0 * n_keep * H, st);
                UNIMM_TRY(gather_rows(nullptr, c.h, keep_rows, n_keep, 2 * H, nullptr, cc.h, st));      // both planes of a row
                UNIMM_TRY(gather_rows(x.f, nullptr, keep_rows, n_keep, H, x_keep->f, nullptr, st));
            }
            // the FFN-1 planes of the kept rows go behind their context planes (n_keep <= M / 2 is not guaranteed: use pre's tail? no —
            // ffn holds [M, 2I] planes; the kept context occupies its first n_keep * 2H elements, which FFN-1 overwrites only
            // after the output projection has consumed them)
            return self_layer_tail(L, cc, *x_keep, pre, ffn, n_keep, st);
        }
        return self_layer_tail(L, c, x, pre, ffn, M, st);
    }
    UNIMM_TRY(linear(x, M, L.qkv, ACT_NONE, nullptr, 0, nullptr, 0, qkv, 3 * H, st));
    const void *qp = qkv, *kp = byte_ptr(qkv) + e * H, *vp = byte_ptr(qkv) + e * 2 * H;
    if (ac.pk != nullptr) {
        const unimm_packed_batch_t& pk = *ac.pk;
        if (text && lp() && D == 64) {
            // tcgen05 path: context jobs and candidate jobs in ONE launch (context rows simply have no own-candidate keys)
            const int n_ctx = pk.n_jobs_text_ctx, n_cand = pk.n_jobs_text_self - pk.n_jobs_text_ctx;
            {
                AttnJobsArgs a;
                a.q = qp; a.ldq = 3 * H; a.k = kp; a.ldk = 3 * H; a.v = vp; a.ldv = 3 * H; a.o = ctx; a.ldo = H;
                a.heads = heads; a.D = D; a.jobs = pk.d_jobs_text_self; a.n_jobs = pk.n_jobs_text_self;
                a.max_q_len = pk.max_q_text_self; a.kv_cap = pk.kv_cap_text; a.win_cap = pk.win_cap;
                a.row_iv = pk.d_row_iv; a.key_mask = nullptr; a.key_mask_ld = 0;
                a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind();
                a.n_rows = M;
                if (attn_umma && attention_candidates_umma_supported(a, pk.cand_halo)) {
                    Prof prof(this, CAT_ATTN, 4.0 * heads * D * pk.pairs_text_self, st);
                    UNIMM_TRY(attention_candidates_umma(a, pk.cand_halo, st));
                    goto attention_done;
                }
            }
            // mma.sync path: context rows attend their context (plain jobs); candidate rows run the persistent double-buffered kernel
            UNIMM_TRY(attention_packed(qp, 3 * H, kp, 3 * H, vp, 3 * H, ctx, H, heads, D, pk.d_jobs_text_self, n_ctx, pk.kv_cap_text,
                                       pk.kv_cap_text, 0, 0.0, ac, st));
            AttnJobsArgs a;
            a.q = qp; a.ldq = 3 * H; a.k = kp; a.ldk = 3 * H; a.v = vp; a.ldv = 3 * H; a.o = ctx; a.ldo = H;
            a.heads = heads; a.D = D; a.jobs = pk.d_jobs_text_self + static_cast<size_t>(n_ctx) * 8; a.n_jobs = n_cand;
            a.max_q_len = pk.max_q_text_self; a.kv_cap = pk.kv_cap_text; a.win_cap = pk.win_cap;
            a.row_iv = pk.d_row_iv; a.key_mask = nullptr; a.key_mask_ld = 0;
            a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind();
            a.n_rows = M;
            Prof prof(this, CAT_ATTN, 4.0 * heads * D * pk.pairs_text_self, st);
            UNIMM_TRY(attention_candidates(a, pk.cand_halo, st));
        } else if (text) UNIMM_TRY(attention_packed(qp, 3 * H, kp, 3 * H, vp, 3 * H, ctx, H, heads, D, pk.d_jobs_text_self, pk.n_jobs_text_self,
                                                    pk.max_q_text_self, pk.kv_cap_text, pk.win_cap, pk.pairs_text_self, ac, st));
        else UNIMM_TRY(attention_packed(qp, 3 * H, kp, 3 * H, vp, 3 * H, ctx, H, heads, D, pk.d_jobs_img_self, pk.n_jobs_img_self,
                                        cfg.num_regions, 64, 0, static_cast<double>(pk.n_units) * cfg.num_regions * cfg.num_regions, ac, st));
    } else {
        const int Sx = text ? cfg.seq_len : cfg.num_regions;
        UNIMM_TRY(attention(qp, 3 * H, kp, 3 * H, vp, 3 * H, ctx, H, ac.B, heads, D, Sx, Sx, text ? MASK_TEXT_SELF : MASK_KEY_VECTOR,
                            text ? ac.desc : nullptr, text ? nullptr : ac.key_mask, st));
    }
attention_done:
    ActBuf c;
    c.f = lp() ? nullptr : static_cast<float*>(ctx); c.h = lp() ? static_cast<bf16*>(ctx) : nullptr; c.ld = H;
    if (keep_rows != nullptr) {
        // only n_keep rows are read downstream: gather their attention context (into the FFN buffer, free until FFN-1 writes it)
        // and their residual stream (into x_keep), and run the row-wise rest of the layer on those rows alone
        UNIMM_CHECK(x_keep != nullptr && n_keep > 0 && n_keep <= M, "self_layer: bad row subset");
        ActBuf cc;
        cc.f = lp() ? nullptr : static_cast<float*>(ffn); cc.h = lp() ? static_cast<bf16*>(ffn) : nullptr; cc.ld = H;
        const bool f32_live = !(lp() && fuse_ln && res16);     // the 16-bit residual mode keeps only x.h current
        {
            Prof prof(this, CAT_ROWWISE, (2.0 * esz() + (f32_live ? 8.0 : 0.0) + (lp() ? 4.0 : 0.0)) * n_keep * H, st);
            UNIMM_TRY(gather_rows(c.f, c.h, keep_rows, n_keep, H, cc.f, cc.h, st));
            UNIMM_TRY(gather_rows(f32_live ? x.f : nullptr, x.h, keep_rows, n_keep, H, f32_live ? x_keep->f : nullptr, lp() ? x_keep->h : nullptr, st));
        }
        return self_layer_tail(L, cc, *x_keep, pre, ffn, n_keep, st);
    }
    return self_layer_tail(L, c, x, pre, ffn, M, st);
}

// out-proj + residual -> LN -> FFN1+GELU -> FFN2 + residual -> LN on rows [0, M) of (attention context c, stream x)
int unimm_engine::self_layer_tail(const SelfLayer& L, const ActBuf& c, ActBuf& x, float* pre, void* ffn, int M, cudaStream_t st) {
    const int H = x.ld;
    UNIMM_TRY(linear_ln(c, M, L.out, x.f, H, L.ln1, pre, x, st));
    const int I = L.ffn1.N;
    UNIMM_TRY(linear(x, M, L.ffn1, ACT_GELU, nullptr, 0, nullptr, 0, ffn, I, st, tc32() ? I : 0));
    ActBuf f;
    const bool f16 = lp() || tc32();            // tc32: the GELU output exists as hi | lo planes only
    f.f = f16 ? nullptr : static_cast<float*>(ffn); f.h = f16 ? static_cast<bf16*>(ffn) : nullptr; f.ld = I;
    UNIMM_TRY(linear_ln(f, M, L.ffn2, x.f, H, L.ln2, pre, x, st));
    return 0;
}

// BertConnectionLayer (reference :770-783): stream 1 = image, stream 2 = text
int unimm_engine::conn_layer(const ConnLayer& L, int Mt, int Mv, const AttnCtx& ac, cudaStream_t st, bool image_out) {
    const unimm_config_t& c = cfg;
    const int S = c.seq_len, R = c.num_regions, H = c.hidden_size, Hv = c.v_hidden_size, Hb = c.bi_hidden_size;
    const int heads = c.bi_num_attention_heads, D = Hb / heads;
    const size_t e = tc32() ? 2 : esz();             // fp32-class mode: Q | K | V live as fp16 planes, hi in [0, 3Hb), lo 3Hb further
    const int pl = tc32() ? 3 * Hb : 0;             // lo-plane offset handed to the projections
    UNIMM_TRY(linear(xv, Mv, L.qkv_v, ACT_NONE, nullptr, 0, nullptr, 0, qkv_v, 3 * Hb, st, pl));
    const int n_sh = (ac.pk != nullptr && kv2_ctx_only) ? ac.pk->n_shared_rows : 0;
    UNIMM_CHECK(image_out || ac.pk != nullptr, "conn_layer: the image update can only be skipped in the packed layout");
    if (!image_out) {
        // nothing reads this layer's image output (scores-only batch, last connection): no image queries, hence no K2 | V2 at all
        Linear q2 = L.qkv_t;
        q2.N = Hb;
        UNIMM_TRY(linear(xt, Mt, q2, ACT_NONE, nullptr, 0, nullptr, 0, qkv_t, 3 * Hb, st, pl));
    } else if (n_sh > 0 && n_sh < Mt && L.qkv_t.w32 != nullptr) {
        // prefix-shared layout: the image rows attend ONLY the context rows (co-mask [1,ctx), utils/data_utils.py:199-210), so the
        // text-side keys / values K2 | V2 (:670-672) of the candidate rows — 86 % of the text rows — are never read: project the
        // queries Q2 for all rows, K2 | V2 for the context rows [0, n_shared) only (the packer puts them first)
        Linear q2 = L.qkv_t, kv2 = L.qkv_t;
        q2.N = Hb;
        kv2.N = 2 * Hb;
        const size_t off = static_cast<size_t>(Hb) * L.qkv_t.K;
        kv2.w32 = L.qkv_t.w32 + off; kv2.b = L.qkv_t.b + Hb;
        if (L.qkv_t.wlp) kv2.wlp = L.qkv_t.wlp + (tc32() ? 2 * off : off);
        if (L.qkv_t.wlp_p16) kv2.wlp_p16 = L.qkv_t.wlp_p16 + off;
        UNIMM_TRY(linear(xt, Mt, q2, ACT_NONE, nullptr, 0, nullptr, 0, qkv_t, 3 * Hb, st, pl));
        UNIMM_TRY(linear(xt, n_sh, kv2, ACT_NONE, nullptr, 0, nullptr, 0, byte_ptr(qkv_t) + e * Hb, 3 * Hb, st, pl));
    } else {
        UNIMM_TRY(linear(xt, Mt, L.qkv_t, ACT_NONE, nullptr, 0, nullptr, 0, qkv_t, 3 * Hb, st, pl));
    }
    if (tc32()) {
        const bf16 *qt = static_cast<const bf16*>(qkv_t), *qv = static_cast<const bf16*>(qkv_v);
        bf16 *ct16 = static_cast<bf16*>(ctx_t), *cv16 = static_cast<bf16*>(ctx_v);
        if (ac.pk != nullptr) {
            const unimm_packed_batch_t& pk = *ac.pk;
            UNIMM_TRY(attention_split(qt, qv + Hb, qv + 2 * Hb, 6 * Hb, 3 * Hb, 6 * Hb, 3 * Hb, ct16, 2 * Hb, Hb, heads, D, pk.d_jobs_t2i, pk.n_jobs_t2i,
                                      pk.max_q_t2i, nullptr, pk.d_image_mask, R, static_cast<double>(Mt) * R, st));
            if (image_out) UNIMM_TRY(attention_split(qv, qt + Hb, qt + 2 * Hb, 6 * Hb, 3 * Hb, 6 * Hb, 3 * Hb, cv16, 2 * Hb, Hb, heads, D, pk.d_jobs_i2t,
                                                     pk.n_jobs_i2t, R, nullptr, nullptr, 0, pk.pairs_i2t, st));
        } else {
            UNIMM_TRY(attention_split(qt, qv + Hb, qv + 2 * Hb, 6 * Hb, 3 * Hb, 6 * Hb, 3 * Hb, ct16, 2 * Hb, Hb, heads, D, dense_jobs, ac.B, S, nullptr,
                                      ac.key_mask, R, static_cast<double>(Mt) * R, st));
            UNIMM_TRY(attention_split(qv, qt + Hb, qt + 2 * Hb, 6 * Hb, 3 * Hb, 6 * Hb, 3 * Hb, cv16, 2 * Hb, Hb, heads, D, dj_i2t, ac.B, R, nullptr,
                                      nullptr, 0, static_cast<double>(Mv) * S, st));
        }
    } else
    if (ac.pk != nullptr) {
        const unimm_packed_batch_t& pk = *ac.pk;
        // text queries (shared + candidate rows of a unit) over the unit's image keys/values (:681-698)
        AttnCtx ax = ac;
        ax.n_q_rows = Mt; ax.n_kv_rows = Mv;
        UNIMM_TRY(attention_packed(qkv_t, 3 * Hb, byte_ptr(qkv_v) + e * Hb, 3 * Hb, byte_ptr(qkv_v) + e * 2 * Hb, 3 * Hb, ctx_t, Hb, heads, D,
                                   pk.d_jobs_t2i, pk.n_jobs_t2i, pk.max_q_t2i, 64, 0, static_cast<double>(Mt) * R, ax, st));
        // image queries over the unit's context rows = the co-attention interval [1,ctx) (:701-721)
        if (image_out) UNIMM_TRY(attention_packed(qkv_v, 3 * Hb, byte_ptr(qkv_t) + e * Hb, 3 * Hb, byte_ptr(qkv_t) + e * 2 * Hb, 3 * Hb, ctx_v, Hb, heads, D,
                                   pk.d_jobs_i2t, pk.n_jobs_i2t, R, pk.kv_cap_text, 0, pk.pairs_i2t, ac, st));
    } else {
        // text queries over image keys/values, image padding mask only (:681-698)
        UNIMM_TRY(attention(qkv_t, 3 * Hb, byte_ptr(qkv_v) + e * Hb, 3 * Hb, byte_ptr(qkv_v) + e * 2 * Hb, 3 * Hb, ctx_t, Hb, ac.B, heads,
                            D, S, R, MASK_KEY_VECTOR, nullptr, ac.key_mask, st));
        // image queries over text keys/values, co-attention column interval only (:701-721)
        UNIMM_TRY(attention(qkv_v, 3 * Hb, byte_ptr(qkv_t) + e * Hb, 3 * Hb, byte_ptr(qkv_t) + e * 2 * Hb, 3 * Hb, ctx_v, Hb, ac.B, heads,
                            D, R, S, MASK_CO_INTERVAL, ac.desc, nullptr, st));
    }
    ActBuf cv, ct;
    const bool c16 = lp() || tc32();                // tc32: the contexts are fp16 hi | lo planes
    cv.f = c16 ? nullptr : static_cast<float*>(ctx_v); cv.h = c16 ? static_cast<bf16*>(ctx_v) : nullptr; cv.ld = Hb;
    ct.f = c16 ? nullptr : static_cast<float*>(ctx_t); ct.h = c16 ? static_cast<bf16*>(ctx_t) : nullptr; ct.ld = Hb;
    // BertBiOutput (:744-754): image rows take the image-query context through dense1, text rows the other through dense2
    if (image_out) UNIMM_TRY(linear_ln(cv, Mv, L.dense1, xv.f, Hv, L.ln1, pre_v, xv, st));
    UNIMM_TRY(linear_ln(ct, Mt, L.dense2, xt.f, H, L.ln2, pre_t, xt, st));
    // image FFN, text FFN (:777-781)
    const int Iv = L.v_ffn1.N, I = L.t_ffn1.N;
    if (image_out) {
        UNIMM_TRY(linear(xv, Mv, L.v_ffn1, ACT_GELU, nullptr, 0, nullptr, 0, ffn_v, Iv, st, tc32() ? Iv : 0));
        ActBuf fv;
        fv.f = (lp() || tc32()) ? nullptr : static_cast<float*>(ffn_v); fv.h = (lp() || tc32()) ? static_cast<bf16*>(ffn_v) : nullptr; fv.ld = Iv;
        UNIMM_TRY(linear_ln(fv, Mv, L.v_ffn2, xv.f, Hv, L.v_ln, pre_v, xv, st));
    }
    UNIMM_TRY(linear(xt, Mt, L.t_ffn1, ACT_GELU, nullptr, 0, nullptr, 0, ffn_t, I, st, tc32() ? I : 0));
    ActBuf ft;
    ft.f = (lp() || tc32()) ? nullptr : static_cast<float*>(ffn_t); ft.h = (lp() || tc32()) ? static_cast<bf16*>(ffn_t) : nullptr; ft.ld = I;
    UNIMM_TRY(linear_ln(ft, Mt, L.t_ffn2, xt.f, H, L.t_ln, pre_t, xt, st));
    return 0;
}

int unimm_engine::forward(const unimm_batch_t& in, const unimm_outputs_t& out, cudaStream_t st) {
    UNIMM_CHECK(finalized, "weights not finalized");
    const unimm_config_t& c = cfg;
    const int B = in.B, S = c.seq_len, R = c.num_regions, H = c.hidden_size, Hv = c.v_hidden_size;
    UNIMM_CHECK(B > 0 && B <= Bmax, "batch size out of range for this engine");
    UNIMM_CHECK(in.d_input_ids && in.d_token_type_ids && in.d_position_ids && in.d_desc && in.d_image_feat && in.d_image_loc &&
                    in.d_image_mask, "required input pointer is NULL");
    const int Mt = B * S, Mv = B * R;
    const SeqDesc* desc = reinterpret_cast<const SeqDesc*>(in.d_desc);

    // ---- embeddings (reference :326-356, :1487-1493)
    UNIMM_TRY(embed_text_ln(in.d_input_ids, in.d_token_type_ids, in.d_position_ids, Mt, H, c.vocab_size, c.max_position_embeddings,
                            c.type_vocab_size, 10, word_emb, pos_emb, type_emb, type_ext_emb, emb_ln.g, emb_ln.b, xt.f, xt.h,
                            act_kind(), err_flag, st));
    UNIMM_TRY(gather_features(in.d_image_feat, in.d_feat_index, B, R, c.v_feature_size, lp() ? nullptr : static_cast<float*>(feat_a),
                              lp() ? static_cast<bf16*>(feat_a) : nullptr, lp_kind(), st));
    UNIMM_TRY(image_loc_embed(in.d_image_loc, in.d_feat_index, B, R, Hv, loc_w, loc_b, pre_v, st));
    {
        ActBuf fa;
        fa.f = lp() ? nullptr : static_cast<float*>(feat_a); fa.h = lp() ? static_cast<bf16*>(feat_a) : nullptr; fa.ld = c.v_feature_size;
        UNIMM_TRY(linear_ln(fa, Mv, img_emb, pre_v, Hv, vemb_ln, pre_v, xv, st));
    }
    // key mask per sequence: image_mask[feat_index[b]] — expand once when an index is used
    const float* key_mask = in.d_image_mask;
    if (in.d_feat_index != nullptr) {
        // reuse vhead as scratch for the expanded [B,R] mask (F = R floats per row; R may not be a multiple of 4)
        float* km = vhead;
        UNIMM_TRY(expand_key_mask(in.d_image_mask, in.d_feat_index, B, R, km, st));
        key_mask = km;
    }

    if (tc32()) UNIMM_TRY(build_dense_jobs(desc, B, S, R, dj_text, dj_i2t, dj_img, dj_row_iv, st));
    // ---- encoder schedule (reference :842-929)
    AttnCtx ac;
    ac.B = B; ac.desc = desc; ac.key_mask = key_mask;
    UNIMM_TRY(run_encoder(Mt, Mv, ac, st));

    if (out.d_sequence_output_t)
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(out.d_sequence_output_t, xt.f, sizeof(float) * Mt * H, cudaMemcpyDeviceToDevice, st));
    if (out.d_sequence_output_v)
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(out.d_sequence_output_v, xv.f, sizeof(float) * Mv * Hv, cudaMemcpyDeviceToDevice, st));

    // ---- poolers + NSP head (reference :946-967, :1062-1070); always fp32
    const bool training = in.d_masked_lm_labels && in.d_next_sentence_label && in.d_image_target;
    float* nsp = out.d_nsp_scores;
    if (nsp == nullptr && training) nsp = label_logit;  // scratch (B*2 <= Mt)
    if (nsp != nullptr)
        UNIMM_TRY(pooler_head(xt.f, H, nullptr, S, xv.f, Hv, nullptr, R, B, nsp, st));
    if (training && out.d_losses)
        UNIMM_TRY(nsp_ce_loss(nsp, in.d_next_sentence_label, B, in.d_nsp_weight, out.d_losses + 2, st));

    // ---- gathered LM head (reference :982-986, :1023-1026 on the labelled rows only) + scores
    const int n = in.n_lm_rows;
    UNIMM_CHECK(n >= 0 && n <= Mt, "n_lm_rows out of range");
    if (n > 0) {
        UNIMM_CHECK(in.d_lm_rows && in.d_masked_lm_labels, "lm rows given without labels");
        UNIMM_TRY(gather_labels(in.d_masked_lm_labels, in.d_lm_rows, n, g_labels, st));
        UNIMM_TRY(lm_head_rows(xt, in.d_lm_rows, g_labels, n, st));
    }
    if (out.d_seq_score || out.d_token_logp || out.d_token_ul)
        UNIMM_TRY(scatter_scores(row_logp, row_ul, in.d_lm_rows, n, B, S, out.d_token_logp, out.d_token_ul, out.d_seq_score, st));

    // ---- losses (reference :1559-1624)
    if (training && out.d_losses) {
        if (in.d_lm_weight) UNIMM_TRY(lm_ul_loss(row_logp, row_ul, in.d_lm_rows, n, in.d_lm_weight, Mt, out.d_losses, st));
        else UNIMM_TRY(lm_ce_loss(row_logp, n, out.d_losses, st));
        // image head (:1085-1088) + masked KL (:1569-1574)
        UNIMM_CHECK(in.d_image_label != nullptr, "image_label required with image_target");
        UNIMM_TRY(linear(xv, Mv, img_transform, ACT_GELU, nullptr, 0, vhead, Hv, nullptr, 0, st));
        { Prof prof(this, CAT_ROWWISE, (8.0 + esz()) * (Mv) * (Hv), st); UNIMM_TRY(layernorm_rows(vhead, Hv, Mv, Hv, img_ln.g, img_ln.b, vhead_h.f, vhead_h.h, act_kind(), st)); }
        UNIMM_TRY(linear(vhead_h, Mv, img_decoder, ACT_NONE, nullptr, 0, v_logits, c.v_target_size, nullptr, 0, st));
        // d_losses[1] = loss, [3..4] scratch
        float* kl = out.d_losses + 3;
        UNIMM_TRY(image_kl_loss(v_logits, c.v_target_size, in.d_image_target, in.d_image_label, Mv, c.v_target_size, kl, st));
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(out.d_losses + 1, kl, sizeof(float), cudaMemcpyDeviceToDevice, st));
    }

    // ---- compatibility: full-vocabulary logits for every text row (reference :1025 as written)
    if (out.d_prediction_scores_t) {
        UNIMM_TRY(linear(xt, Mt, lm_transform, ACT_GELU, nullptr, 0, pre_t, H, nullptr, 0, st));
        ActBuf hh;
        hh.f = pre_t; hh.h = lp() ? static_cast<bf16*>(ffn_t) : nullptr; hh.ld = H;
        { Prof prof(this, CAT_ROWWISE, (8.0 + esz()) * (Mt) * (H), st); UNIMM_TRY(layernorm_rows(pre_t, H, Mt, H, lm_ln.g, lm_ln.b, hh.f, hh.h, ln_kind(), st)); }
        UNIMM_TRY(linear(hh, Mt, lm_decoder, ACT_NONE, nullptr, 0, out.d_prediction_scores_t, c.vocab_size, nullptr, 0, st));
    }
    return 0;
}

// encoder layer schedule (reference BertEncoder.forward :842-929) over whichever row layout `ac` describes
int unimm_engine::run_encoder(int Mt, int Mv, const AttnCtx& ac, cudaStream_t st) {
    const unimm_config_t& c = cfg;
    int v_start = 0, t_start = 0;
    auto run_v = [&](int i) { return self_layer(v_layers[i], xv, pre_v, qkv_v, ctx_v, ffn_v, Mv, c.v_num_attention_heads, false, ac, st); };
    auto run_t = [&](int i) { return self_layer(t_layers[i], xt, pre_t, qkv_t, ctx_t, ffn_t, Mt, c.num_attention_heads, true, ac, st); };
    // Scores-only packed batch (no [CLS] rows, only the labelled rows' log-likelihoods are asked for): after the last connection
    // the image stream feeds nothing but the pooled NSP logit (:946-967), and of the last text layer's output only the labelled
    // rows reach the LM head — the other rows matter to that layer as keys / values alone.
    const bool lean = prune_tail && ac.pk != nullptr && ac.pk->no_cls_rows && ac.pk->n_lm_rows > 0 && c.num_connections > 0;
    xk_live = false;
    for (int k = 0; k < c.num_connections; ++k) {
        for (int i = v_start; i < c.v_biattention_id[k]; ++i) UNIMM_TRY(run_v(i));
        for (int i = t_start; i < c.t_biattention_id[k]; ++i) UNIMM_TRY(run_t(i));
        UNIMM_TRY(conn_layer(c_layers[k], Mt, Mv, ac, st, !(lean && k == c.num_connections - 1)));
        v_start = c.v_biattention_id[k];
        t_start = c.t_biattention_id[k];
    }
    if (!lean) {
        for (int i = v_start; i < c.v_num_hidden_layers; ++i) UNIMM_TRY(run_v(i));
        for (int i = t_start; i < c.num_hidden_layers; ++i) UNIMM_TRY(run_t(i));
    } else {
        for (int i = t_start; i < c.num_hidden_layers - 1; ++i) UNIMM_TRY(run_t(i));
        if (t_start < c.num_hidden_layers) {
            UNIMM_TRY(self_layer(t_layers[c.num_hidden_layers - 1], xt, pre_t, qkv_t, ctx_t, ffn_t, Mt, c.num_attention_heads, true, ac, st,
                                 keep_rows_, n_keep_, &xk));
            xk_live = true;
        }
        return 0;     // nobody reads the fp32 views of a scores-only batch
    }
    if (lp() && fuse_ln && res16) {   // the poolers and the optional sequence outputs read the fp32 view
        Prof prof(this, CAT_ROWWISE, 6.0 * (static_cast<double>(Mt) * xt.ld + static_cast<double>(Mv) * xv.ld), st);
        UNIMM_TRY(cast_lp_to_f32(xt.h, xt.f, static_cast<size_t>(Mt) * xt.ld, ln_kind(), st));
        UNIMM_TRY(cast_lp_to_f32(xv.h, xv.f, static_cast<size_t>(Mv) * xv.ld, ln_kind(), st));
    }
    return 0;
}

// gathered LM head (reference :982-986, :1023-1026 on the labelled rows only): fills row_logp / row_ul [n]
// transform + LayerNorm of the LM head (reference :982-986) on n rows of src (gathered through d_rows when given): g_h.f and the
// operand of the vocabulary GEMM (16-bit rows, or fp16 hi | lo planes in the fp32-class mode and in the bf16 mode's lm_hp head)
int unimm_engine::lm_transform_rows(const ActBuf& src, const int* d_rows, int n, cudaStream_t st) {
    const int H = cfg.hidden_size;
    if (lp() && lm_hp) {
        UNIMM_CHECK(src.f != nullptr, "fp32-class LM head: the fp32 residual stream is not live");
        { Prof prof(this, CAT_ROWWISE, 8.0 * n * H, st); UNIMM_TRY(split_f32_to_hilo(src.f, src.ld, n, H, lm_split, st, d_rows)); }
        GemmEpilogue ep;
        ep.bias = lm_transform.b; ep.act = ACT_GELU_ERF; ep.lp_kind = LP_FP16; ep.split3 = 1; ep.out_f32 = g_t1; ep.ldo_f32 = H;
        { Prof prof(this, CAT_GEMM, 2.0 * n * H * H, st); UNIMM_TRY(gemm_umma_bf16(lm_split, 2 * H, lm_transform.whl, 2 * H, n, H, H, ep, 0, 0, st)); }
        Prof prof(this, CAT_ROWWISE, 12.0 * n * H, st);
        return layernorm_rows(g_t1, H, n, H, lm_ln.g, lm_ln.b, g_h.f, g_h_hl, LP_HILO, st);
    }
    if (d_rows != nullptr)
        UNIMM_TRY(gather_rows(lp() ? nullptr : src.f, lp() ? src.h : nullptr, d_rows, n, H, lp() ? nullptr : g_in.f, lp() ? g_in.h : nullptr, st));
    UNIMM_TRY(linear(d_rows != nullptr ? g_in : src, n, lm_transform, ACT_GELU, nullptr, 0, g_t1, H, nullptr, 0, st));
    Prof prof(this, CAT_ROWWISE, (8.0 + esz()) * (n) * (H), st);
    return layernorm_rows(g_t1, H, n, H, lm_ln.g, lm_ln.b, g_h.f, g_h.h, act_kind(), st);
}

// gathered LM head (reference :982-986, :1023-1026 on the labelled rows only): fills row_logp / row_ul [n]
int unimm_engine::lm_head_rows(const ActBuf& src, const int* d_rows, const int* d_labels, int n, cudaStream_t st) {
    const unimm_config_t& c = cfg;
    const int H = c.hidden_size;
    UNIMM_TRY(lm_transform_rows(src, d_rows, n, st));
    if (lp() || tc32()) {
        // the vocabulary GEMM with the online log-sum-exp epilogue: logits never leave TMEM / registers (fp32-class mode and lm_hp:
        // the same epilogue over the three-pass accumulator)
        GemmEpilogue ep;
        ep.bias = lm_decoder.b;
        ep.labels = d_labels;
        ep.partials = partials;
        ep.label_logit = label_logit;
        ep.lp_kind = lm32() ? LP_FP16 : ln_kind();
        ep.split3 = lm32() ? 1 : 0;
        const int ld = lm32() ? 2 * H : H;
        const bf16* a = lm32() && !tc32() ? g_h_hl : g_h.h;
        Prof prof(this, CAT_LMHEAD, 2.0 * n * c.vocab_size * H, st);
        UNIMM_TRY(gemm_umma_bf16(a, ld, lm32() ? planes_of(lm_decoder) : lm_decoder.wlp, ld, n, c.vocab_size, H, ep, 256, 0, st));
        UNIMM_TRY(lse_from_partials(partials, gemm_umma_lse_tiles(c.vocab_size), label_logit, n, row_logp, row_ul, st));
    } else {
        for (int r0 = 0; r0 < n; r0 += kLogitRows) {
            const int nr = std::min(kLogitRows, n - r0);
            GemmEpilogue ep;
            ep.bias = lm_decoder.b;
            ep.out_f32 = logits_chunk;
            ep.ldo_f32 = logits_ld();
            Prof prof(this, CAT_LMHEAD, 2.0 * nr * c.vocab_size * H, st);
            UNIMM_TRY(gemm_simt_f32(g_h.f + static_cast<size_t>(r0) * H, H, lm_decoder.w32, H, nr, c.vocab_size, H, ep, st));
            UNIMM_TRY(lse_from_logits(logits_chunk, logits_ld(), nr, c.vocab_size, d_labels + r0, row_logp + r0, row_ul + r0, st));
        }
    }
    return 0;
}

// LM head with shared labelled rows (16-bit modes): transform + LayerNorm + vocabulary log-sum-exp once per unique row, then the
// label logits of the n entries as gathered dot products
int unimm_engine::lm_head_shared(const ActBuf& src, const int* d_urows, int n_u, const int* d_uidx, const int* d_labels, int n,
                                 cudaStream_t st) {
    const unimm_config_t& c = cfg;
    const int H = c.hidden_size;
    UNIMM_CHECK(lp() || tc32(), "shared labelled rows need a tensor-core mode");
    UNIMM_TRY(lm_transform_rows(src, d_urows, n_u, st));
    UNIMM_CUDA_CHECK(cudaMemsetAsync(g_labels, 0, sizeof(int) * n_u, st));      // the fused epilogue's label pick is unused here
    {
        GemmEpilogue ep;
        ep.bias = lm_decoder.b;
        ep.labels = g_labels;
        ep.partials = partials;
        ep.label_logit = label_logit;
        ep.lp_kind = lm32() ? LP_FP16 : ln_kind();
        ep.split3 = lm32() ? 1 : 0;
        const int ld = lm32() ? 2 * H : H;
        const bf16* a = lm32() && !tc32() ? g_h_hl : g_h.h;
        Prof prof(this, CAT_LMHEAD, 2.0 * n_u * c.vocab_size * H, st);
        UNIMM_TRY(gemm_umma_bf16(a, ld, lm32() ? planes_of(lm_decoder) : lm_decoder.wlp, ld, n_u, c.vocab_size, H, ep, 256, 0, st));
    }
    UNIMM_TRY(lse_merge(partials, gemm_umma_lse_tiles(c.vocab_size), n_u, lse_u, st));
    if (lm32()) UNIMM_TRY(label_scores_f32(g_h.f, H, lm_decoder.w32, H, lm_decoder.b, d_uidx, d_labels, lse_u, n, H, row_logp, row_ul, st));
    else UNIMM_TRY(label_scores(g_h.h, H, lm_decoder.wlp, H, lm_decoder.b, d_uidx, d_labels, lse_u, n, H, ln_kind(), row_logp, row_ul, st));
    return 0;
}

// Prefix-shared generative scoring over packed rows (see attention_jobs.cu and unimm_b200/packing.py)
int unimm_engine::forward_packed(const unimm_packed_batch_t& in, float* d_seq_score, float* d_nsp_scores, float* d_token_logp,
                                 cudaStream_t st) {
    UNIMM_CHECK(finalized, "weights not finalized");
    const unimm_config_t& c = cfg;
    const int U = in.n_units, C = in.n_cands, M = in.n_text_rows, R = c.num_regions, H = c.hidden_size, Hv = c.v_hidden_size;
    UNIMM_CHECK(U > 0 && C > 0 && M > 0, "empty packed batch");
    UNIMM_CHECK(static_cast<long long>(M) <= static_cast<long long>(Bmax) * c.seq_len && U <= Bmax, "packed batch exceeds the engine workspace");
    UNIMM_CHECK(in.kv_cap_text > 0 && in.kv_cap_text <= 256 && in.kv_cap_text % 64 == 0 && in.win_cap % 64 == 0, "bad staging capacities");
    UNIMM_CHECK(in.n_lm_rows >= 0 && in.n_lm_rows <= M, "n_lm_rows out of range");
    UNIMM_CHECK(!(d_nsp_scores && in.no_cls_rows), "NSP scores requested from a batch packed without [CLS] rows (scores_only)");
    const int Mv = U * R;
    // the job kernels stage the image keys of a unit in a 64-row buffer
    UNIMM_CHECK(R <= 64, "the prefix-shared layout supports at most 64 image regions per unit");
    UNIMM_CHECK(in.win_cap >= 128 + 2 * in.cand_halo, "win_cap too small for the longest candidate (needs 128 + 2 * cand_halo)");
    UNIMM_TRY(embed_text_ln_i32(in.d_input_ids, in.d_token_type_ids, in.d_position_ids, M, H, c.vocab_size, c.max_position_embeddings,
                                c.type_vocab_size, 10, word_emb, pos_emb, type_emb, type_ext_emb, emb_ln.g, emb_ln.b, xt.f, xt.h, act_kind(),
                                err_flag, st));
    UNIMM_CHECK(in.d_unit_image == nullptr || in.n_images > 0, "d_unit_image given without n_images");
    UNIMM_TRY(gather_features(in.d_image_feat, in.d_unit_image, U, R, c.v_feature_size, lp() ? nullptr : static_cast<float*>(feat_a),
                              lp() ? static_cast<bf16*>(feat_a) : nullptr, lp_kind(), st));
    UNIMM_TRY(image_loc_embed(in.d_image_loc, in.d_unit_image, U, R, Hv, loc_w, loc_b, pre_v, st));
    {
        ActBuf fa;
        fa.f = lp() ? nullptr : static_cast<float*>(feat_a); fa.h = lp() ? static_cast<bf16*>(feat_a) : nullptr; fa.ld = c.v_feature_size;
        UNIMM_TRY(linear_ln(fa, Mv, img_emb, pre_v, Hv, vemb_ln, pre_v, xv, st));
    }
    AttnCtx ac;
    ac.pk = &in;
    const int n = in.n_lm_rows;
    // labelled rows shared by several candidates (the unit-wide B_0 row of a scores-only batch): LM head once per unique row
    const bool dedup = (lp() || tc32()) && lm_dedup && n > 0 && in.n_lm_unique > 0 && in.n_lm_unique < n && in.d_lm_urows && in.d_lm_uidx;
    keep_rows_ = dedup ? in.d_lm_urows : in.d_lm_rows;
    n_keep_ = dedup ? in.n_lm_unique : n;
    UNIMM_TRY(run_encoder(M, Mv, ac, st));
    if (d_nsp_scores)
        UNIMM_TRY(pooler_head(xt.f, H, in.d_cand_cls_row, 0, xv.f, Hv, in.d_cand_img_row, 0, C, d_nsp_scores, st));
    if (n > 0) {
        if (dedup) UNIMM_TRY(lm_head_shared(xk_live ? xk : xt, xk_live ? nullptr : in.d_lm_urows, in.n_lm_unique, in.d_lm_uidx, in.d_lm_labels, n, st));
        else if (xk_live) UNIMM_TRY(lm_head_rows(xk, nullptr, in.d_lm_labels, n, st));       // the last text layer already gathered them
        else UNIMM_TRY(lm_head_rows(xt, in.d_lm_rows, in.d_lm_labels, n, st));
    }
    if (d_seq_score) UNIMM_TRY(segment_sum(row_logp, in.d_cand_lm_off, C, d_seq_score, st));
    if (d_token_logp && n > 0) UNIMM_CUDA_CHECK(cudaMemcpyAsync(d_token_logp, row_logp, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// =================================================================================================
// C ABI
// =================================================================================================
namespace {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

extern "C" {

const char* unimm_last_error(void) { return unimm::get_error(); }
int unimm_abi_version(void) { return UNIMM_ABI_VERSION; }
int64_t unimm_launch_count(void) { return g_launches.load(); }
void unimm_reset_launch_count(void) { g_launches.store(0); }

int unimm_create(const unimm_config_t* cfg, int device, int precision, int max_sequences, unimm_engine_t** out) {
    UNIMM_CHECK(cfg != nullptr && out != nullptr, "null argument");
    UNIMM_CHECK(precision == UNIMM_PREC_FP32 || precision == UNIMM_PREC_BF16 || precision == UNIMM_PREC_FP16, "unknown precision");
    UNIMM_CHECK(max_sequences > 0, "max_sequences must be positive");
    UNIMM_CHECK(cfg->hidden_size == 768 || cfg->hidden_size == 1024, "hidden_size must be 768 or 1024");
    UNIMM_CHECK(cfg->v_hidden_size == 768 || cfg->v_hidden_size == 1024, "v_hidden_size must be 768 or 1024");
    UNIMM_CHECK(cfg->num_connections >= 0 && cfg->num_connections <= 16, "too many connection layers");
    const int d = cfg->hidden_size / cfg->num_attention_heads, dv = cfg->v_hidden_size / cfg->v_num_attention_heads,
              db = cfg->bi_hidden_size / cfg->bi_num_attention_heads;
    UNIMM_CHECK((d == 64 || d == 128) && (dv == 64 || dv == 128) && (db == 64 || db == 128), "head dims must be 64 or 128");
    UNIMM_CHECK(cfg->seq_len > 0 && cfg->seq_len <= 256 && cfg->num_regions > 0 && cfg->num_regions <= 256,
                "seq_len / num_regions must be in (0, 256]");
    int ndev = 0;
    UNIMM_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    UNIMM_CHECK(device >= 0 && device < ndev, "no such CUDA device");
    cudaDeviceProp prop;
    UNIMM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    UNIMM_CHECK(prop.major == 10, "unimm_b200 kernels are built for sm_100a (B200) only");
    DeviceGuard g(device);
    unimm_engine* e = new unimm_engine();
    e->cfg = *cfg;
    e->device = device;
    e->prec = precision;
    e->Bmax = max_sequences;
    if (const char* f = getenv("UNIMM_FUSE_LN")) e->fuse_ln = atoi(f) != 0;   // A/B switches for bench.py; defaults: on
    if (const char* f = getenv("UNIMM_FRAG_EPILOGUE")) e->frag_epilogue = atoi(f) != 0;
    // fp16 mode: the 1-SFU form of the SAME erf-GELU (error <= |x| * 2.4e-4, half an fp16 ulp of the stored value; config-1 sequence
    // log-likelihoods move by < 1e-3) — the FFN-1 epilogue is SFU-bound otherwise.  bf16 is already marginal against its bound: exact form.
    e->gelu_tanh = precision == UNIMM_PREC_FP16;
    if (const char* f = getenv("UNIMM_GELU_TANH")) e->gelu_tanh = atoi(f) != 0;
    if (const char* f = getenv("UNIMM_ATTN_UMMA")) e->attn_umma = atoi(f) != 0;
    if (const char* f = getenv("UNIMM_KV2_ALL")) e->kv2_ctx_only = atoi(f) == 0;
    if (const char* f = getenv("UNIMM_PRUNE_TAIL")) e->prune_tail = atoi(f) != 0;
    if (const char* f = getenv("UNIMM_LM_DEDUP")) e->lm_dedup = atoi(f) != 0;
    e->tc32_ = precision == UNIMM_PREC_FP32;
    if (const char* f = getenv("UNIMM_FP32_SIMT")) e->tc32_ = e->tc32_ && atoi(f) == 0;
    e->mix16 = precision == UNIMM_PREC_BF16;
    if (const char* f = getenv("UNIMM_BF16_PURE")) e->mix16 = e->mix16 && atoi(f) == 0;
    e->lm_hp = false;
    if (const char* f = getenv("UNIMM_LM_HP")) e->lm_hp = atoi(f) != 0 && e->lp();
    // 16-bit residual stream: the residual of a LayerNorm-fused projection is the previous LayerNorm's own 16-bit output, added on the
    // tensor core.  That copy is fp16 in the fp16 mode AND in the bf16 mode's mixed formats (11 significant bits either way); with bf16
    // for every operand it would be an 8-bit residual stream, and the fp32-class LM head wants the fp32 stream: both keep fp32.
    e->res16 = e->fuse_ln && (precision == UNIMM_PREC_FP16 || (precision == UNIMM_PREC_BF16 && e->mix16)) && !e->lm_hp;
    if (const char* f = getenv("UNIMM_RES16")) e->res16 = e->res16 && atoi(f) != 0;
    *out = e;
    return 0;
}

int unimm_destroy(unimm_engine_t* e) {
    if (e == nullptr) return 0;
    DeviceGuard g(e->device);
    cudaDeviceSynchronize();
    if (e->host_path.h_rows) cudaFreeHost(e->host_path.h_rows);
    if (e->h_err) cudaFreeHost(e->h_err);
    for (cudaEvent_t ev : e->slot_done)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->h2d_done)
        if (ev) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (void* p : e->owned) cudaFree(p);
    delete e;
    return 0;
}

int unimm_load_weight(unimm_engine_t* e, const char* name, const float* h_data, const int64_t* shape, int ndim) {
    UNIMM_CHECK(e && name && h_data && shape && ndim > 0 && ndim <= 4, "bad argument");
    UNIMM_CHECK(!e->finalized, "weights already finalized");
    DeviceGuard g(e->device);
    std::string key(name);
    const std::string prefix = "bert_pretrained.";
    if (key.compare(0, prefix.size(), prefix) == 0) key = key.substr(prefix.size());
    DevTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) {
        t.shape.push_back(shape[i]);
        t.numel *= static_cast<size_t>(shape[i]);
    }
    UNIMM_CHECK(t.numel > 0, "empty tensor");
    // keep every buffer 16-byte sized so 128-bit casts/gathers never read past the end
    UNIMM_TRY(e->dalloc(&t.p, (t.numel + 3) & ~size_t(3)));
    UNIMM_CUDA_CHECK(cudaMemcpy(t.p, h_data, t.numel * sizeof(float), cudaMemcpyHostToDevice));
    e->raw[key] = t;
    return 0;
}

int unimm_finalize_weights(unimm_engine_t* e) {
    UNIMM_CHECK(e != nullptr, "null engine");
    DeviceGuard g(e->device);
    return e->finalize();
}

int unimm_forward(unimm_engine_t* e, const unimm_batch_t* batch, const unimm_outputs_t* out, void* stream) {
    UNIMM_CHECK(e && batch && out, "null argument");
    DeviceGuard g(e->device);
    return e->forward(*batch, *out, static_cast<cudaStream_t>(stream));
}

int unimm_forward_packed(unimm_engine_t* e, const unimm_packed_batch_t* batch, float* d_seq_score, float* d_nsp_scores,
                         float* d_token_logp, void* stream) {
    UNIMM_CHECK(e && batch, "null argument");
    DeviceGuard g(e->device);
    return e->forward_packed(*batch, d_seq_score, d_nsp_scores, d_token_logp, static_cast<cudaStream_t>(stream));
}

int unimm_submit_packed_host(unimm_engine_t* e, const unimm_packed_batch_t* hb, int slot, float* h_seq_score, float* h_nsp_scores, void* stream) {
    UNIMM_CHECK(e && hb && h_seq_score, "null argument");
    UNIMM_CHECK(slot == 0 || slot == 1, "slot must be 0 or 1");
    UNIMM_CHECK(e->finalized, "weights not finalized");
    const unimm_config_t& c = e->cfg;
    const int U = hb->n_units, C = hb->n_cands, M = hb->n_text_rows, R = c.num_regions, F = c.v_feature_size;
    UNIMM_CHECK(U > 0 && C > 0 && M > 0 && static_cast<long long>(M) <= static_cast<long long>(e->Bmax) * c.seq_len && U <= e->Bmax &&
                    C <= e->Bmax * c.seq_len, "packed batch exceeds the engine workspace");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PackedStage& ps = e->packed_stage[slot];
    if (e->slot_done[slot] == nullptr) UNIMM_CUDA_CHECK(cudaEventCreateWithFlags(&e->slot_done[slot], cudaEventDisableTiming));
    if (e->h2d_done[slot] == nullptr) UNIMM_CUDA_CHECK(cudaEventCreateWithFlags(&e->h2d_done[slot], cudaEventDisableTiming));
    if (e->copy_stream == nullptr) UNIMM_CUDA_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    // The slot's previous occupant has been waited for (unimm_wait_packed) before it is reused, so its staging arena is free NOW:
    // upload on the copy stream, concurrently with whatever the other slot's step is still computing on `st`.
    cudaStream_t cs = e->copy_stream;
    // (a caller that reuses a slot WITHOUT having waited still gets ordered behind the previous occupant: no-op once it is done)
    if (e->slot_used[slot]) UNIMM_CUDA_CHECK(cudaStreamWaitEvent(cs, e->slot_done[slot], 0));
    e->slot_used[slot] = true;
    if (ps.i32 == nullptr) {
        const size_t rows = static_cast<size_t>(e->Bmax) * c.seq_len;
        // ids, types, pos (3M) + row_iv (4M) + lm rows/labels/unique rows/indices (4M) + cand arrays (3C+1 <= 3M+1) + jobs (6U*8)
        ps.i32_cap = rows * 14 + static_cast<size_t>(e->Bmax) * 52 + 64;
        UNIMM_TRY(e->dalloc(&ps.i32, ps.i32_cap));
        ps.f32_cap = static_cast<size_t>(e->Bmax) * R * (F + 6) + rows * 3 + 64;
        UNIMM_TRY(e->dalloc(&ps.f32, ps.f32_cap));
        UNIMM_CUDA_CHECK(cudaDeviceSynchronize());      // dalloc clears on the legacy stream, which the non-blocking copy stream does not follow
    }
    // the staged key ranges must fit the capacities the kernels were sized with (the arrays are host memory here: check them)
    for (int j = 0; j < hb->n_jobs_text_self; ++j)
        UNIMM_CHECK(hb->d_jobs_text_self[8 * j + 3] <= hb->kv_cap_text, "a text job's key range exceeds kv_cap_text");
    for (int j = 0; j < hb->n_jobs_i2t; ++j)
        UNIMM_CHECK(hb->d_jobs_i2t[8 * j + 3] <= hb->kv_cap_text, "an image->text job's key range exceeds kv_cap_text");
    for (int j = 0; j < hb->n_jobs_t2i; ++j) UNIMM_CHECK(hb->d_jobs_t2i[8 * j + 3] <= 64, "a text->image job has more than 64 keys");
    // carve the device staging buffers and copy each host array behind the previous one
    unimm_packed_batch_t d = *hb;
    size_t io = 0, fo = 0;
    auto put_i = [&](const int32_t* h, size_t n, const int32_t** out) -> int {
        *out = nullptr;
        if (h == nullptr || n == 0) return 0;
        UNIMM_CHECK(io + n <= ps.i32_cap, "packed host batch larger than the staging buffer");
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(ps.i32 + io, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
        *out = ps.i32 + io;
        io += (n + 3) & ~size_t(3);
        return 0;
    };
    auto put_f = [&](const float* h, size_t n, const float** out) -> int {
        UNIMM_CHECK(h != nullptr && fo + n <= ps.f32_cap, "packed host batch larger than the staging buffer");
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(ps.f32 + fo, h, n * sizeof(float), cudaMemcpyHostToDevice, cs));
        *out = ps.f32 + fo;
        fo += (n + 3) & ~size_t(3);
        return 0;
    };
    UNIMM_TRY(put_i(hb->d_input_ids, M, &d.d_input_ids));
    UNIMM_TRY(put_i(hb->d_token_type_ids, M, &d.d_token_type_ids));
    UNIMM_TRY(put_i(hb->d_position_ids, M, &d.d_position_ids));
    UNIMM_TRY(put_i(hb->d_row_iv, static_cast<size_t>(M) * 4, &d.d_row_iv));
    UNIMM_TRY(put_i(hb->d_jobs_text_self, static_cast<size_t>(hb->n_jobs_text_self) * 8, &d.d_jobs_text_self));
    UNIMM_TRY(put_i(hb->d_jobs_t2i, static_cast<size_t>(hb->n_jobs_t2i) * 8, &d.d_jobs_t2i));
    UNIMM_TRY(put_i(hb->d_jobs_i2t, static_cast<size_t>(hb->n_jobs_i2t) * 8, &d.d_jobs_i2t));
    UNIMM_TRY(put_i(hb->d_jobs_img_self, static_cast<size_t>(hb->n_jobs_img_self) * 8, &d.d_jobs_img_self));
    UNIMM_TRY(put_i(hb->d_lm_rows, hb->n_lm_rows, &d.d_lm_rows));
    UNIMM_TRY(put_i(hb->d_lm_labels, hb->n_lm_rows, &d.d_lm_labels));
    UNIMM_TRY(put_i(hb->d_cand_lm_off, static_cast<size_t>(C) + 1, &d.d_cand_lm_off));
    UNIMM_TRY(put_i(hb->d_cand_cls_row, C, &d.d_cand_cls_row));
    UNIMM_TRY(put_i(hb->d_cand_img_row, C, &d.d_cand_img_row));
    UNIMM_TRY(put_i(hb->d_lm_urows, hb->n_lm_unique, &d.d_lm_urows));
    UNIMM_TRY(put_i(hb->d_lm_uidx, hb->d_lm_urows != nullptr && hb->n_lm_unique > 0 ? hb->n_lm_rows : 0, &d.d_lm_uidx));
    const int NI = hb->d_unit_image != nullptr ? hb->n_images : U;      // feature blocks: one per image, or one per unit
    UNIMM_CHECK(NI > 0 && NI <= U, "n_images out of range");
    UNIMM_TRY(put_i(hb->d_unit_image, hb->d_unit_image != nullptr ? U : 0, &d.d_unit_image));
    UNIMM_TRY(put_f(hb->d_image_feat, static_cast<size_t>(NI) * R * F, &d.d_image_feat));
    UNIMM_TRY(put_f(hb->d_image_loc, static_cast<size_t>(NI) * R * 5, &d.d_image_loc));
    UNIMM_TRY(put_f(hb->d_image_mask, static_cast<size_t>(NI) * R, &d.d_image_mask));
    float* d_score = ps.f32 + fo;
    float* d_nsp = d_score + ((C + 3) & ~3);
    UNIMM_CHECK(fo + static_cast<size_t>(C) * 3 + 8 <= ps.f32_cap, "packed host batch larger than the staging buffer");
    UNIMM_CUDA_CHECK(cudaEventRecord(e->h2d_done[slot], cs));
    UNIMM_CUDA_CHECK(cudaStreamWaitEvent(st, e->h2d_done[slot], 0));
    UNIMM_TRY(e->forward_packed(d, d_score, h_nsp_scores ? d_nsp : nullptr, nullptr, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(h_seq_score, d_score, sizeof(float) * C, cudaMemcpyDeviceToHost, st));
    if (h_nsp_scores) UNIMM_CUDA_CHECK(cudaMemcpyAsync(h_nsp_scores, d_nsp, sizeof(float) * 2 * C, cudaMemcpyDeviceToHost, st));
    UNIMM_TRY(e->queue_id_check(st, slot));
    UNIMM_CUDA_CHECK(cudaEventRecord(e->slot_done[slot], st));
    return 0;
}

int unimm_wait_packed(unimm_engine_t* e, int slot) {
    UNIMM_CHECK(e != nullptr && (slot == 0 || slot == 1), "bad argument");
    UNIMM_CHECK(e->slot_done[slot] != nullptr, "nothing was submitted to this slot");
    DeviceGuard g(e->device);
    UNIMM_CUDA_CHECK(cudaEventSynchronize(e->slot_done[slot]));
    return e->id_check_result(slot);
}

int unimm_score_packed_host(unimm_engine_t* e, const unimm_packed_batch_t* hb, float* h_seq_score, float* h_nsp_scores, void* stream) {
    UNIMM_TRY(unimm_submit_packed_host(e, hb, 0, h_seq_score, h_nsp_scores, stream));
    return unimm_wait_packed(e, 0);
}

int unimm_verify_masks(const unimm_seq_desc_t* d_desc, int B, int S, int R, const void* d_txt_mask, int txt_elem_bytes,
                       const int64_t* d_co_mask, int* d_mismatch, void* stream) {
    UNIMM_CHECK(d_desc && d_txt_mask && d_mismatch && B > 0, "bad argument");
    return verify_masks(reinterpret_cast<const SeqDesc*>(d_desc), B, S, R, d_txt_mask, txt_elem_bytes, 0, d_co_mask, d_mismatch,
                        static_cast<cudaStream_t>(stream));
}

int unimm_score_host(unimm_engine_t* e, const unimm_host_batch_t* hb, float* h_seq_score, float* h_nsp_scores, void* stream) {
    UNIMM_CHECK(e && hb && h_seq_score, "null argument");
    UNIMM_CHECK(e->finalized, "weights not finalized");
    const unimm_config_t& c = e->cfg;
    const int B = hb->B, U = hb->U, S = c.seq_len, R = c.num_regions, F = c.v_feature_size;
    UNIMM_CHECK(B > 0 && B <= e->Bmax && U > 0 && U <= e->Bmax, "batch size out of range for this engine");
    UNIMM_CHECK(hb->h_feat_index != nullptr || U == B, "feat_index required when U != B");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    HostPath& hp = e->host_path;
    if (hp.ids == nullptr) {
        const size_t n = static_cast<size_t>(e->Bmax);
        UNIMM_TRY(e->dalloc(&hp.ids, n * S)); UNIMM_TRY(e->dalloc(&hp.types, n * S)); UNIMM_TRY(e->dalloc(&hp.pos, n * S));
        UNIMM_TRY(e->dalloc(&hp.labels, n * S)); UNIMM_TRY(e->dalloc(&hp.desc, n));
        UNIMM_TRY(e->dalloc(&hp.feat, n * R * F)); UNIMM_TRY(e->dalloc(&hp.loc, n * R * 5)); UNIMM_TRY(e->dalloc(&hp.mask, n * R));
        UNIMM_TRY(e->dalloc(&hp.index, n)); UNIMM_TRY(e->dalloc(&hp.rows, n * S));
        UNIMM_TRY(e->dalloc(&hp.score, n)); UNIMM_TRY(e->dalloc(&hp.nsp, 2 * n));
        UNIMM_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&hp.h_rows), sizeof(int32_t) * n * S));
    }
    // labelled positions (label != -1), in order: the rows the LM head runs on
    int n_rows = 0;
    if (hb->h_masked_lm_labels != nullptr) {
        const int64_t* lab = hb->h_masked_lm_labels;
        for (int i = 0; i < B * S; ++i)
            if (lab[i] != -1) hp.h_rows[n_rows++] = i;
    }
    const size_t bs = static_cast<size_t>(B) * S;
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.ids, hb->h_input_ids, bs * 8, cudaMemcpyHostToDevice, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.types, hb->h_token_type_ids, bs * 8, cudaMemcpyHostToDevice, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.pos, hb->h_position_ids, bs * 8, cudaMemcpyHostToDevice, st));
    if (n_rows > 0) {
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.labels, hb->h_masked_lm_labels, bs * 8, cudaMemcpyHostToDevice, st));
        UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.rows, hp.h_rows, sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice, st));
    }
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.desc, hb->h_desc, sizeof(unimm_seq_desc_t) * B, cudaMemcpyHostToDevice, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.feat, hb->h_image_feat, sizeof(float) * U * R * F, cudaMemcpyHostToDevice, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.loc, hb->h_image_loc, sizeof(float) * U * R * 5, cudaMemcpyHostToDevice, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.mask, hb->h_image_mask, sizeof(float) * U * R, cudaMemcpyHostToDevice, st));
    if (hb->h_feat_index) UNIMM_CUDA_CHECK(cudaMemcpyAsync(hp.index, hb->h_feat_index, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));

    unimm_batch_t in;
    std::memset(&in, 0, sizeof(in));
    in.B = B;
    in.d_input_ids = hp.ids; in.d_token_type_ids = hp.types; in.d_position_ids = hp.pos;
    in.d_desc = hp.desc; in.d_image_feat = hp.feat; in.d_image_loc = hp.loc; in.d_image_mask = hp.mask;
    in.d_feat_index = hb->h_feat_index ? hp.index : nullptr;
    in.d_masked_lm_labels = n_rows > 0 ? hp.labels : nullptr;
    in.d_lm_rows = hp.rows;
    in.n_lm_rows = n_rows;
    unimm_outputs_t out;
    std::memset(&out, 0, sizeof(out));
    out.d_seq_score = hp.score;
    out.d_nsp_scores = h_nsp_scores ? hp.nsp : nullptr;
    UNIMM_TRY(e->forward(in, out, st));
    UNIMM_CUDA_CHECK(cudaMemcpyAsync(h_seq_score, hp.score, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
    if (h_nsp_scores) UNIMM_CUDA_CHECK(cudaMemcpyAsync(h_nsp_scores, hp.nsp, sizeof(float) * 2 * B, cudaMemcpyDeviceToHost, st));
    UNIMM_TRY(e->queue_id_check(st));
    UNIMM_CUDA_CHECK(cudaStreamSynchronize(st));
    return e->id_check_result();
}

int unimm_check_ids(unimm_engine_t* e, void* stream) {
    UNIMM_CHECK(e != nullptr && e->finalized, "null or unfinalized engine");
    DeviceGuard g(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    UNIMM_TRY(e->queue_id_check(st));
    UNIMM_CUDA_CHECK(cudaStreamSynchronize(st));
    return e->id_check_result();
}

int unimm_rank_metrics(const float* d_scores, int rows, int n_opt, const int32_t* d_gt_index, const float* d_relevance,
                       int32_t* d_ranks, double* d_sums, void* stream) {
    UNIMM_CHECK(d_scores != nullptr, "null scores");
    return rank_metrics(d_scores, rows, n_opt, d_gt_index, d_relevance, d_ranks, d_sums, static_cast<cudaStream_t>(stream));
}

int unimm_neural_ndcg(const float* d_y_pred, const float* d_y_true, int rows, int n_opt, float temperature, int max_iter, float tol,
                      float* d_ndcg, float* d_idcg, void* stream) {
    UNIMM_CHECK(d_y_pred && d_y_true && d_ndcg && d_idcg, "null argument");
    return neural_ndcg(d_y_pred, d_y_true, rows, n_opt, temperature, max_iter, tol, d_ndcg, d_idcg, static_cast<cudaStream_t>(stream));
}

int unimm_ensemble_normalise(const float* d_probs, int n_models, int rows, int n_opt, float* d_out, void* stream) {
    UNIMM_CHECK(d_probs && d_out, "null argument");
    return ensemble_normalise(d_probs, n_models, rows, n_opt, d_out, static_cast<cudaStream_t>(stream));
}

int unimm_profile_begin(unimm_engine_t* e) {
    UNIMM_CHECK(e != nullptr, "null engine");
    e->profiling = true;
    return 0;
}

// Synchronises the device, then sums per class: elapsed ms, work units (FLOPs, or bytes for CAT_ROWWISE), launches.
int unimm_profile_end(unimm_engine_t* e, double* ms, double* work, int64_t* launches, int ncat) {
    UNIMM_CHECK(e && ms && work && launches && ncat >= unimm_engine::NCAT, "bad argument");
    DeviceGuard g(e->device);
    e->profiling = false;
    UNIMM_CUDA_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < ncat; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
    for (double& b : e->prof_bytes) b = 0;
    for (auto& r : e->prof_recs) {
        float t = 0.f;
        UNIMM_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cat] += t; work[r.cat] += r.work; launches[r.cat] += 1;
        e->prof_bytes[r.cat] += r.bytes;
        e->prof_pool.push_back(r.a); e->prof_pool.push_back(r.b);
    }
    e->prof_recs.clear();
    return 0;
}

int unimm_profile_bytes(unimm_engine_t* e, double* bytes, int ncat) {
    UNIMM_CHECK(e && bytes && ncat >= unimm_engine::NCAT, "bad argument");
    for (int i = 0; i < ncat; ++i) bytes[i] = i < unimm_engine::NCAT ? e->prof_bytes[i] : 0.0;
    return 0;
}

// ---------------------------------------------------------------------------------- kernel-level entry points
int unimm_k_gemm_lp(const void* d_A, int lda, const void* d_W, int ldw, int M, int N, int K, const float* d_bias,
                    const float* d_residual, int ldr, int act, float* d_out_f32, int ldo_f32, void* d_out_lp, int ldo_lp,
                    int tile_n, int max_ctas, int lp_kind, void* stream) {
    GemmEpilogue ep;
    ep.w_perm16 = (lp_kind & 0x100) != 0;
    lp_kind &= 0xff;
    ep.lp_kind = lp_kind;
    ep.debug_mode = tile_n / 1000;   // microbenchmark hook (scripts/gemm_bench.py): tile_n = 1000*mode + tile
    tile_n %= 1000;
    ep.pre_act_f32 = (act & 0x300) != 0;      // act | 0x100: d_out_f32 receives the pre-activation, d_out_lp the activation (training forward)
    ep.pre_act_lp = (act & 0x200) != 0;       // act | 0x200: the same, d_out_f32 pointing at a 16-bit [M, ldo_f32] matrix (lp_kind's encoding)
    act &= 0xff;
    ep.bias = d_bias; ep.residual = d_residual; ep.ldr = ldr; ep.act = act;
    ep.out_f32 = d_out_f32; ep.ldo_f32 = ldo_f32; ep.out_bf16 = static_cast<bf16*>(d_out_lp); ep.ldo_bf16 = ldo_lp;
    return gemm_umma_bf16(static_cast<const bf16*>(d_A), lda, static_cast<const bf16*>(d_W), ldw, M, N, K, ep, tile_n, max_ctas,
                          static_cast<cudaStream_t>(stream));
}

int unimm_k_permute_w_ln(const void* d_W_lp, void* d_Wp_lp, int N, int K, void* stream) {
    return permute_weight_rows_ln(static_cast<const bf16*>(d_W_lp), static_cast<bf16*>(d_Wp_lp), N, K, static_cast<cudaStream_t>(stream));
}
int unimm_k_permute_w(const void* d_W_lp, void* d_Wp_lp, int N, int K, int mode, void* stream) {
    return permute_weight_rows(static_cast<const bf16*>(d_W_lp), static_cast<bf16*>(d_Wp_lp), N, K, mode, static_cast<cudaStream_t>(stream));
}

int unimm_k_gemm_ln_lp(const void* d_A, int lda, const void* d_W, int ldw, int M, int N, int K, const float* d_bias,
                       const float* d_residual, int ldr, const void* d_residual_lp, int ldr_lp, const float* d_gamma, const float* d_beta,
                       float* d_out_f32, int ldo_f32, void* d_out_lp, int ldo_lp, int lp_kind, void* stream) {
    GemmLnEpilogue ep;
    ep.bias = d_bias; ep.residual = d_residual; ep.ldr = ldr; ep.gamma = d_gamma; ep.beta = d_beta;
    ep.residual_lp = static_cast<const bf16*>(d_residual_lp); ep.ldr_lp = ldr_lp;
    ep.out_f32 = d_out_f32; ep.ldo_f32 = ldo_f32; ep.out_lp = static_cast<bf16*>(d_out_lp); ep.ldo_lp = ldo_lp; ep.lp_kind = lp_kind;
    return gemm_umma_ln(static_cast<const bf16*>(d_A), lda, static_cast<const bf16*>(d_W), ldw, M, N, K, ep,
                        static_cast<cudaStream_t>(stream));
}

int unimm_k_gemm_f32(const float* d_A, int lda, const float* d_W, int ldw, int M, int N, int K, const float* d_bias,
                     const float* d_residual, int ldr, int act, float* d_out_f32, int ldo_f32, void* stream) {
    GemmEpilogue ep;
    ep.bias = d_bias; ep.residual = d_residual; ep.ldr = ldr; ep.act = act;
    ep.out_f32 = d_out_f32; ep.ldo_f32 = ldo_f32;
    return gemm_simt_f32(d_A, lda, d_W, ldw, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

int unimm_k_lm_head_lp(const void* d_H, int ldh, const void* d_E, int lde, int rows, int V, int K, const float* d_bias,
                       const int32_t* d_labels, float* d_partials_scratch, float* d_label_logit_scratch, float* d_logp,
                       float* d_ul, int lp_kind, void* stream) {
    GemmEpilogue ep;
    ep.lp_kind = lp_kind;
    ep.bias = d_bias;
    ep.labels = d_labels;
    ep.partials = reinterpret_cast<float2*>(d_partials_scratch);
    ep.label_logit = d_label_logit_scratch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    UNIMM_TRY(gemm_umma_bf16(static_cast<const bf16*>(d_H), ldh, static_cast<const bf16*>(d_E), lde, rows, V, K, ep, 256, 0, st));
    return lse_from_partials(ep.partials, gemm_umma_lse_tiles(V), d_label_logit_scratch, rows, d_logp, d_ul, st);
}

size_t unimm_k_lm_head_backward_scratch(int rows, int V, int K) {
    const size_t Vp = (static_cast<size_t>(V) + 63) / 64 * 64, np = (static_cast<size_t>(rows) + 63) / 64 * 64;
    // dz [rows, Vp] + dzT [V, np] + Et [K, Vp] + hT [K, np] 16-bit, partials float2 [rows, tiles], lse / coef / label logit / logp / ul fp32 [rows]
    return 2 * (rows * Vp + V * np + K * Vp + K * np) + 8 * static_cast<size_t>(rows) * gemm_umma_lse_tiles(V) + 4 * 5 * static_cast<size_t>(rows) + 4096;
}

int unimm_k_lm_head_backward(const void* d_H, int ldh, const void* d_E, int lde, int rows, int V, int K, const float* d_bias,
                             const int32_t* d_labels, const float* d_weight, float grad_scale, float* d_dH, float* d_dE, float* d_dbias,
                             float* d_logp, void* d_scratch, size_t scratch_bytes, int lp_kind, void* stream) {
    UNIMM_CHECK(d_H && d_E && d_bias && d_labels && d_weight && d_dH && d_dE && d_scratch && rows > 0 && V > 0 && K % 64 == 0, "bad argument");
    UNIMM_CHECK(scratch_bytes >= unimm_k_lm_head_backward_scratch(rows, V, K), "scratch smaller than unimm_k_lm_head_backward_scratch()");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Vp = (V + 63) / 64 * 64, np = (rows + 63) / 64 * 64, tiles = gemm_umma_lse_tiles(V);
    char* p = static_cast<char*>(d_scratch);
    auto carve = [&](size_t bytes) { char* q = p; p += (bytes + 255) & ~size_t(255); return q; };
    bf16* dz = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(rows) * Vp));
    bf16* dzT = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(V) * np));
    bf16* Et = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(K) * Vp));
    bf16* hT = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(K) * np));
    float2* partials = reinterpret_cast<float2*>(carve(8 * static_cast<size_t>(rows) * tiles));
    float* lse = reinterpret_cast<float*>(carve(4 * static_cast<size_t>(rows)));
    float* coef = reinterpret_cast<float*>(carve(4 * static_cast<size_t>(rows)));
    float* lab_logit = reinterpret_cast<float*>(carve(4 * static_cast<size_t>(rows)));
    float* logp = reinterpret_cast<float*>(carve(4 * static_cast<size_t>(rows)));
    float* ul = reinterpret_cast<float*>(carve(4 * static_cast<size_t>(rows)));
    UNIMM_CHECK(static_cast<size_t>(p - static_cast<char*>(d_scratch)) <= scratch_bytes, "scratch carve overflow");
    // 1. forward recompute: log-sum-exp per row (and log p of the label) without materialising logits
    {
        GemmEpilogue ep;
        ep.lp_kind = lp_kind; ep.bias = d_bias; ep.labels = d_labels; ep.partials = partials; ep.label_logit = lab_logit;
        UNIMM_TRY(gemm_umma_bf16(static_cast<const bf16*>(d_H), ldh, static_cast<const bf16*>(d_E), lde, rows, V, K, ep, 256, 0, st));
        UNIMM_TRY(lse_merge(partials, tiles, rows, lse, st));
        UNIMM_TRY(lse_from_partials(partials, tiles, lab_logit, rows, logp, ul, st));
        if (d_logp) UNIMM_CUDA_CHECK(cudaMemcpyAsync(d_logp, logp, sizeof(float) * rows, cudaMemcpyDeviceToDevice, st));
    }
    // 2. dz = coef (softmax - onehot), both orientations, 16-bit.  The factor kScale keeps the ~1/V-sized probabilities out of the
    // fp16 subnormal range; the two gradient GEMMs multiply it back out (alpha)
    const float kScale = lp_kind == LP_FP16 ? 1024.f : 1.f;
    UNIMM_TRY(lm_loss_coef(logp, d_weight, rows, grad_scale * kScale, coef, st));
    UNIMM_CUDA_CHECK(cudaMemsetAsync(dz, 0, 2 * static_cast<size_t>(rows) * Vp, st));
    UNIMM_CUDA_CHECK(cudaMemsetAsync(dzT, 0, 2 * static_cast<size_t>(V) * np, st));
    {
        GemmEpilogue ep;
        ep.lp_kind = lp_kind; ep.bias = d_bias; ep.labels = d_labels; ep.lse = lse; ep.coef = coef;
        ep.dz = dz; ep.ldz = Vp; ep.dz_cols = Vp; ep.dzT = dzT; ep.ldzt = np;
        UNIMM_TRY(gemm_umma_bf16(static_cast<const bf16*>(d_H), ldh, static_cast<const bf16*>(d_E), lde, rows, V, K, ep, 256, 0, st));
    }
    // 3. dH = dz E  (contraction over the vocabulary: W operand = E^T, K-major)
    UNIMM_CUDA_CHECK(cudaMemsetAsync(Et, 0, 2 * static_cast<size_t>(K) * Vp, st));
    UNIMM_TRY(transpose_16(static_cast<const bf16*>(d_E), lde, V, K, Et, Vp, st));
    {
        GemmEpilogue ep;
        ep.lp_kind = lp_kind; ep.out_f32 = d_dH; ep.ldo_f32 = K; ep.alpha = 1.f / kScale;
        UNIMM_TRY(gemm_umma_bf16(dz, Vp, Et, Vp, rows, K, Vp, ep, 0, 0, st));
    }
    // 4. dE = dz^T h  (contraction over the rows: A = dz^T, W = h^T)
    UNIMM_CUDA_CHECK(cudaMemsetAsync(hT, 0, 2 * static_cast<size_t>(K) * np, st));
    UNIMM_TRY(transpose_16(static_cast<const bf16*>(d_H), ldh, rows, K, hT, np, st));
    {
        GemmEpilogue ep;
        ep.lp_kind = lp_kind; ep.out_f32 = d_dE; ep.ldo_f32 = K; ep.alpha = 1.f / kScale;
        UNIMM_TRY(gemm_umma_bf16(dzT, np, hT, np, V, K, np, ep, 0, 0, st));
    }
    // 5. dbias = column sums of dz = row sums of dz^T
    if (d_dbias) UNIMM_TRY(row_sums_16(dzT, np, V, rows, lp_kind, 1.f / kScale, d_dbias, st));
    return 0;
}

size_t unimm_k_linear_backward_scratch(int M, int N, int K) {
    const size_t Mp = (static_cast<size_t>(M) + 63) / 64 * 64;
    // dY16 [M, N] + dY16^T [N, Mp] + X^T [K, Mp] + W^T [K, N] (16-bit) + the scale pair
    return 2 * (static_cast<size_t>(M) * N + N * Mp + K * Mp + static_cast<size_t>(K) * N) + 4096;
}

int unimm_k_linear_backward(const float* d_dY, int ldy, const void* d_X, int ldx, const void* d_W, int ldw, int M, int N, int K, float* d_dX,
                            float* d_dW, float* d_db, void* d_scratch, size_t scratch_bytes, int lp_kind, void* stream) {
    return unimm_k_linear_backward_acc(d_dY, ldy, d_X, ldx, d_W, ldw, M, N, K, d_dX, 0, d_dW, d_db, nullptr, nullptr, nullptr, 0u, 0.f, d_scratch,
                                       scratch_bytes, lp_kind, stream);
}

int unimm_k_linear_backward_acc(const float* d_dY, int ldy, const void* d_X, int ldx, const void* d_W, int ldw, int M, int N, int K, float* d_dX,
                                int accumulate_dx, float* d_dW, float* d_db, const float* d_amax, const float* d_gelu_t, float* d_dX_amax,
                                uint32_t drop_seed, float drop_p, void* d_scratch, size_t scratch_bytes, int lp_kind, void* stream) {
    return unimm_k_linear_backward_phase(d_dY, ldy, d_X, ldx, d_W, ldw, M, N, K, d_dX, accumulate_dx, d_dW, d_db, d_amax, d_gelu_t, d_dX_amax, drop_seed,
                                         drop_p, d_scratch, scratch_bytes, lp_kind, 0, stream);
}

// phase 0: everything on `stream`.  phase 1: the pass over dY (16-bit copy + scale into the scratch, bias gradient) and the dgrad GEMM.
// phase 2: the wgrad GEMM alone, from the scratch a phase-1 call with the same shape filled — so that a caller can run it on a second
// stream beside whatever follows the dgrad (unimm_b200/train_ops.py: the HBM-bound LayerNorm / cast passes hide under it).
int unimm_k_linear_backward_phase(const float* d_dY, int ldy, const void* d_X, int ldx, const void* d_W, int ldw, int M, int N, int K, float* d_dX,
                                  int accumulate_dx, float* d_dW, float* d_db, const float* d_amax, const float* d_gelu_t, float* d_dX_amax,
                                  uint32_t drop_seed, float drop_p, void* d_scratch, size_t scratch_bytes, int lp_kind, int phase, void* stream) {
    UNIMM_CHECK(phase >= 0 && phase <= 2, "phase: 0 = all, 1 = cast + dgrad, 2 = wgrad from the scratch");
    const int gelu_t_kind = (lp_kind & 0x100) != 0 ? (lp_kind & 0xff) : -1;     // lp_kind | 0x100: d_gelu_t points at 16-bit values
    lp_kind &= 0xff;
    if (phase == 1) d_dW = nullptr;
    if (phase == 2) { d_dX = nullptr; d_db = nullptr; }
    UNIMM_CHECK(d_dY && d_X && d_W && d_scratch && M > 0 && N > 0 && K > 0, "bad argument");
    UNIMM_CHECK(N % 64 == 0 && K % 8 == 0 && ldy % 2 == 0, "linear backward: N must be a multiple of 64 (the dgrad contraction), K of 8");
    UNIMM_CHECK(scratch_bytes >= unimm_k_linear_backward_scratch(M, N, K), "scratch smaller than unimm_k_linear_backward_scratch()");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Mp = (M + 63) / 64 * 64;
    char* p = static_cast<char*>(d_scratch);
    auto carve = [&](size_t bytes) { char* q = p; p += (bytes + 255) & ~size_t(255); return q; };
    float* scale = reinterpret_cast<float*>(carve(16));
    bf16* dY16 = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(M) * N));
    bf16* dYT = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(N) * Mp));
    bf16* XT = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(K) * Mp));
    bf16* WT = reinterpret_cast<bf16*>(carve(2 * static_cast<size_t>(K) * N));
    UNIMM_CHECK(static_cast<size_t>(p - static_cast<char*>(d_scratch)) <= scratch_bytes, "scratch carve overflow");
    // the incoming gradient as a 16-bit operand: fp16 gets a power-of-two scale from its own maximum (gradients are routinely below
    // fp16's normal range), multiplied back out by the GEMM epilogues straight from device memory
    // ... and in the same pass over dY: the bias gradient (column sums) and, when dY still has to pass the erf GELU of this projection's
    // output backwards (d_gelu_t = the saved pre-activation), that derivative (|gelu'| <= 1.13 bounds the scaled values)
    if (phase != 2) {
        UNIMM_TRY(amax_scale(d_dY, static_cast<size_t>(M) * ldy, lp_kind == LP_FP16 ? 1 : 0, scale, st, d_amax, d_gelu_t != nullptr ? 1.13f : 1.f));
        UNIMM_CHECK(drop_p <= 0.f || (d_gelu_t == nullptr && static_cast<double>(M) * N < 4294967296.0), "output dropout: not in front of a GELU; 32-bit index");
        UNIMM_TRY(cast_colsum_lp(d_dY, ldy, d_gelu_t, N, M, N, scale, dY16, N, lp_kind, d_db, st, make_drop(drop_seed, drop_p), gelu_t_kind));
    }
    d_db = nullptr;                                              // done
    static const bool transposed_copies = getenv("UNIMM_BWD_TRANSPOSE") != nullptr && atoi(getenv("UNIMM_BWD_TRANSPOSE")) != 0;
    UNIMM_CHECK(phase == 0 || !transposed_copies, "UNIMM_BWD_TRANSPOSE (the first version's transposed copies) runs as one phase");
    if (!transposed_copies) {
        // No transposed copies: tcgen05 reads an operand whose contraction index is the ROW of the stored matrix as an MN-major tile
        // (gemm_umma.cu, a_mn / b_mn).  dX [M, K] = dY [M, N] . W [N, K]: W as stored is the B operand [K, N-contraction] MN-major.
        if (d_dX != nullptr) {
            GemmEpilogue ep;
            ep.lp_kind = lp_kind; ep.out_f32 = d_dX; ep.ldo_f32 = K; ep.alpha_ptr = scale + 1; ep.b_mn = true;
            if (accumulate_dx) { ep.residual = d_dX; ep.ldr = K; }   // dX += dY W: each element is read and rewritten by the same thread
            if (d_dX_amax != nullptr) {                              // max |dX| for whoever turns dX into a 16-bit operand next
                UNIMM_CUDA_CHECK(cudaMemsetAsync(d_dX_amax, 0, sizeof(float), st));
                ep.amax_out = reinterpret_cast<unsigned*>(d_dX_amax);
            }
            UNIMM_TRY(gemm_umma_bf16(dY16, N, static_cast<const bf16*>(d_W), ldw, M, K, N, ep, 0, 0, st));
        }
        // dW [N, K] = dY^T [N, M] . X [M, K]: both operands as stored (contraction = their rows, zero-filled beyond M by TMA).  The
        // output is at most 3072 x 3072 — a handful of tiles — and the contraction tens of thousands of rows long: split it over
        // enough CTAs to fill the machine, partial products added with atomics onto the zeroed gradient.
        if (d_dW != nullptr) {
            const int tn = (K % 256 == 0 || K > 2048) ? 256 : 128;
            const int tiles = ((N + 127) / 128) * ((K + tn - 1) / tn), num_k = (M + 63) / 64;
            int sms = 148;
            { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
            int best = 1;
            double best_eff = 0.0;
            for (int sk = 1; sk <= 16 && sk <= num_k; ++sk) {
                const int total = tiles * sk;
                const double eff = static_cast<double>(total) / (((total + sms - 1) / sms) * sms);
                if (eff > best_eff + 0.02) { best_eff = eff; best = sk; }
            }
            GemmEpilogue ep;
            ep.lp_kind = lp_kind; ep.out_f32 = d_dW; ep.ldo_f32 = K; ep.alpha_ptr = scale + 1; ep.a_mn = true; ep.b_mn = true; ep.split_k = best;
            if (best > 1) UNIMM_CUDA_CHECK(cudaMemsetAsync(d_dW, 0, sizeof(float) * static_cast<size_t>(N) * K, st));
            UNIMM_TRY(gemm_umma_bf16(dY16, N, static_cast<const bf16*>(d_X), ldx, N, K, M, ep, tn, 0, st));
        }
        if (d_db != nullptr) UNIMM_TRY(column_sums_f32(d_dY, ldy, M, N, d_db, st));
        return 0;
    }
    if (d_dX != nullptr) {        // dX [M, K] = dY [M, N] · W [N, K]: contraction over N, W operand = W^T (K-major in N)
        UNIMM_TRY(transpose_16(static_cast<const bf16*>(d_W), ldw, N, K, WT, N, st));
        GemmEpilogue ep;
        ep.lp_kind = lp_kind; ep.out_f32 = d_dX; ep.ldo_f32 = K; ep.alpha_ptr = scale + 1;
        if (accumulate_dx) { ep.residual = d_dX; ep.ldr = K; }       // dX += dY W: each element is read and rewritten by the same thread
        if (d_dX_amax != nullptr) {
            UNIMM_CUDA_CHECK(cudaMemsetAsync(d_dX_amax, 0, sizeof(float), st));
            ep.amax_out = reinterpret_cast<unsigned*>(d_dX_amax);
        }
        UNIMM_TRY(gemm_umma_bf16(dY16, N, WT, N, M, K, N, ep, 0, 0, st));
    }
    if (d_dW != nullptr) {        // dW [N, K] = dY^T [N, M] · X [M, K]: contraction over the rows (zero-padded to a multiple of 64)
        UNIMM_CUDA_CHECK(cudaMemsetAsync(dYT, 0, 2 * static_cast<size_t>(N) * Mp, st));
        UNIMM_CUDA_CHECK(cudaMemsetAsync(XT, 0, 2 * static_cast<size_t>(K) * Mp, st));
        UNIMM_TRY(transpose_16(dY16, N, M, N, dYT, Mp, st));
        UNIMM_TRY(transpose_16(static_cast<const bf16*>(d_X), ldx, M, K, XT, Mp, st));
        GemmEpilogue ep;
        ep.lp_kind = lp_kind; ep.out_f32 = d_dW; ep.ldo_f32 = K; ep.alpha_ptr = scale + 1;
        UNIMM_TRY(gemm_umma_bf16(dYT, Mp, XT, Mp, N, K, Mp, ep, 0, 0, st));
    }
    if (d_db != nullptr) UNIMM_TRY(column_sums_f32(d_dY, ldy, M, N, d_db, st));
    return 0;
}

int unimm_k_layernorm(const float* d_x, int ldx, int rows, int H, const float* d_gamma, const float* d_beta, float* d_y_f32,
                      void* d_y_lp, int lp_kind, void* stream) {
    return layernorm_rows(d_x, ldx, rows, H, d_gamma, d_beta, d_y_f32, static_cast<bf16*>(d_y_lp), lp_kind, static_cast<cudaStream_t>(stream));
}

int unimm_k_layernorm_backward(const float* d_dy, const float* d_x, int rows, int H, const float* d_gamma, float* d_dx, float* d_dgamma,
                               float* d_dbeta, void* stream) {
    UNIMM_CHECK(d_dy && d_x && d_gamma && d_dx && d_dgamma && d_dbeta, "null argument");
    return layernorm_backward(d_dy, d_x, rows, H, d_gamma, d_dx, d_dgamma, d_dbeta, static_cast<cudaStream_t>(stream));
}

int unimm_k_gelu_backward(const float* d_dy, const float* d_x, int64_t n, float* d_dx, void* stream) {
    UNIMM_CHECK(d_dy && d_x && d_dx && n > 0, "bad argument");
    return gelu_backward(d_dy, d_x, static_cast<size_t>(n), d_dx, static_cast<cudaStream_t>(stream));
}

int unimm_k_cast_lp(const float* d_src, void* d_dst, int64_t n, int lp_kind, void* stream) {
    return cast_f32_to_lp(d_src, static_cast<bf16*>(d_dst), static_cast<size_t>(n), lp_kind, static_cast<cudaStream_t>(stream));
}

int unimm_k_attention_jobs(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo, int n_rows,
                           int heads, int D, const int32_t* d_jobs, int n_jobs, int max_q_len, int kv_cap, int win_cap,
                           const int32_t* d_row_iv, int halo, int lp_kind, int impl, void* stream) {
    AttnJobsArgs a;
    a.q = d_q; a.ldq = ldq; a.k = d_k; a.ldk = ldk; a.v = d_v; a.ldv = ldv; a.o = d_o; a.ldo = ldo;
    a.heads = heads; a.D = D; a.jobs = d_jobs; a.n_jobs = n_jobs; a.max_q_len = max_q_len; a.kv_cap = kv_cap; a.win_cap = win_cap;
    a.row_iv = d_row_iv; a.key_mask = nullptr; a.key_mask_ld = 0;
    a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind; a.n_rows = n_rows;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == 0) return attention_jobs(a, false, st);
    if (impl == 1) return attention_candidates(a, halo, st);
    UNIMM_CHECK(impl == 2, "impl: 0 = generic jobs, 1 = persistent mma.sync candidates, 2 = tcgen05 candidates");
    return attention_candidates_umma(a, halo, st);
}

int unimm_k_attention_cross_jobs(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo,
                                 int n_q_rows, int n_kv_rows, int heads, int D, const int32_t* d_jobs, int n_jobs, int max_q_len,
                                 const float* d_key_mask, int key_mask_ld, int lp_kind, int impl, void* stream) {
    AttnJobsArgs a;
    a.q = d_q; a.ldq = ldq; a.k = d_k; a.ldk = ldk; a.v = d_v; a.ldv = ldv; a.o = d_o; a.ldo = ldo;
    a.heads = heads; a.D = D; a.jobs = d_jobs; a.n_jobs = n_jobs; a.max_q_len = max_q_len; a.kv_cap = 64; a.win_cap = 0;
    a.row_iv = nullptr; a.key_mask = d_key_mask; a.key_mask_ld = key_mask_ld;
    a.scale = 1.0f / sqrtf(static_cast<float>(D)); a.lp_kind = lp_kind; a.n_rows = n_q_rows; a.n_kv_rows = n_kv_rows;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (impl == 0) return attention_jobs(a, false, st);
    UNIMM_CHECK(impl == 2, "impl: 0 = generic jobs, 2 = tcgen05 cross attention");
    return attention_cross_umma(a, st);
}

int unimm_k_attention(const void* d_q, int ldq, const void* d_k, int ldk, const void* d_v, int ldv, void* d_o, int ldo, int B,
                      int heads, int D, int Sq, int Skv, int mask_kind, const unimm_seq_desc_t* d_desc, const float* d_key_mask,
                      int elem_kind, int impl, void* stream) {
    AttnArgs a;
    a.q = d_q; a.ldq = ldq; a.k = d_k; a.ldk = ldk; a.v = d_v; a.ldv = ldv; a.o = d_o; a.ldo = ldo;
    a.B = B; a.heads = heads; a.D = D; a.Sq = Sq; a.Skv = Skv; a.mask_kind = mask_kind;
    a.desc = reinterpret_cast<const SeqDesc*>(d_desc); a.key_mask = d_key_mask;
    a.scale = 1.0f / sqrtf(static_cast<float>(D));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (elem_kind == 0) {
        UNIMM_CHECK(impl == 0, "the tensor-core attention takes 16-bit tensors");
        return attention_simt_f32(a, st);
    }
    a.lp_kind = elem_kind == 2 ? LP_FP16 : LP_BF16;
    if (impl == 2) {          // tcgen05 kernel of the dense layout (text self-attention only)
        UNIMM_CHECK(mask_kind == MASK_TEXT_SELF && Sq == Skv, "impl 2 is the dense text self-attention");
        AttnJobsArgs j;
        j.q = d_q; j.ldq = ldq; j.k = d_k; j.ldk = ldk; j.v = d_v; j.ldv = ldv; j.o = d_o; j.ldo = ldo;
        j.heads = heads; j.D = D; j.jobs = nullptr; j.n_jobs = B; j.max_q_len = Sq; j.kv_cap = 256; j.win_cap = 0;
        j.row_iv = nullptr; j.key_mask = nullptr; j.key_mask_ld = 0; j.scale = a.scale; j.lp_kind = a.lp_kind;
        j.n_rows = B * Sq; j.desc = a.desc; j.seq_len = Sq;
        return attention_dense_umma(j, st);
    }
    return impl == 0 ? attention_simt_lp(a, st) : attention_mma_lp(a, st);
}

}  // extern "C"
