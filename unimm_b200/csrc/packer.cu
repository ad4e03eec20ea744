// Host-side packer: the reference's per-image [rows, S] int64 tensors (dataloader_visdial.py:437-457, as val_lm.py:55-121
// flattens them) -> the prefix-shared layout of unimm_packed_batch_t (include/unimm_b200.h), written straight into pinned
// staging buffers.  Same layout, row for row, as the vectorised numpy packer in unimm_b200/packing.py (which stays as the
// readable specification and is required to agree in tests/test_packer_cpu.py); this one is what the sweep and the bench call:
// plain loops over units on a few worker threads, ~1 ms for a step of 80 units / 8 000 candidates instead of ~20 ms.
//
// Row layout: [all units' context rows (dense columns 1 .. ctx-1 of candidate 0) | per unit: its shared B_0 row (if any),
// then per candidate: [CLS] (unless scores_only), A_0 .. A_{na-1}, B_{b0} .. B_{nb-1}].  With S positions per sequence a
// candidate truncated at S (utils/data_utils.py:205-209, :237-244) simply has fewer B (and, beyond that, A) rows:
//   nb = clamp(S - L, 0, last)          masked-copy rows that exist
//   na = min(last, S - ctx)             visible-copy rows that exist; scores_only keeps A_0 .. A_{nb-2} (all B_k can see)
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/unimm_b200.h"
#include "common.cuh"

namespace {

constexpr int Q_TILE = 128;   // query rows per CTA of the candidate attention (attention_jobs.cu / attention_umma.cu)

struct UnitPlan {
    int ctx = 0, sh_len = 0, b0_shared = 0;
    int own_rows = 0;        // candidate rows without the shared B_0 row
    int n_lm = 0, n_lm_unique = 0;
    int max_rep = 0;
    long long pairs = 0;
    // prefix sums (filled serially)
    int sh_start = 0, base = 0, cand0 = 0, lm0 = 0, ulm0 = 0;
};

template <typename F>
void parallel_units(int n, int threads, F&& fn) {
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto& th : pool) th.join();
}

inline int roundup(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct unimm_packer {
    int S = 0, R = 0, F = 0;
    bool pinned = false;
    int32_t* ibuf = nullptr;
    size_t icap = 0;
    bool ipinned = false;
    float* fbuf = nullptr;
    size_t fcap = 0;
    bool fpinned = false;
    std::vector<unimm_seq_desc_t> desc;      // per candidate, unit order
    std::vector<int> cand_rep, cand_na, cand_nb;
    std::vector<UnitPlan> plan;
    std::vector<std::string> unit_err;
    unimm_packed_batch_t batch;
    bool valid = false;

    int grow(void** p, size_t* cap, bool* is_pinned, size_t need, size_t elem) {
        if (need <= *cap) return 0;
        const size_t n = need + need / 4 + 1024;
        void* q = nullptr;
        bool pin = pinned;
        if (pin && cudaMallocHost(&q, n * elem) != cudaSuccess) {
            cudaGetLastError();
            q = nullptr;
            pin = false;                              // no pinned memory left: plain pages still work (slower H2D)
        }
        if (q == nullptr) q = malloc(n * elem);
        UNIMM_CHECK(q != nullptr, "packer: out of host memory");
        release(*p, *is_pinned);
        *p = q;
        *cap = n;
        *is_pinned = pin;
        return 0;
    }
    static void release(void* p, bool is_pinned) {
        if (p == nullptr) return;
        if (is_pinned) cudaFreeHost(p);
        else free(p);
    }
};

extern "C" {

int unimm_packer_create(int seq_len, int num_regions, int feature_size, int pinned, unimm_packer_t** out) {
    UNIMM_CHECK(out != nullptr && seq_len > 1 && seq_len <= 256 && num_regions > 0 && feature_size > 0, "packer: bad dimensions");
    unimm_packer* p = new unimm_packer();
    p->S = seq_len; p->R = num_regions; p->F = feature_size;
    if (pinned) {
        int n = 0;
        p->pinned = cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
        if (!p->pinned) cudaGetLastError();
    }
    std::memset(&p->batch, 0, sizeof(p->batch));
    *out = p;
    return 0;
}

int unimm_packer_destroy(unimm_packer_t* p) {
    if (p == nullptr) return 0;
    unimm_packer::release(p->ibuf, p->ipinned);
    unimm_packer::release(p->fbuf, p->fpinned);
    delete p;
    return 0;
}

int unimm_packer_batch(const unimm_packer_t* p, unimm_packed_batch_t* out) {
    UNIMM_CHECK(p != nullptr && out != nullptr, "null argument");
    UNIMM_CHECK(p->valid, "packer holds no batch (the last unimm_packer_pack failed or none was made)");
    *out = p->batch;
    return 0;
}

int unimm_packer_desc(const unimm_packer_t* p, const unimm_seq_desc_t** out, int32_t* n) {
    UNIMM_CHECK(p != nullptr && out != nullptr && n != nullptr, "null argument");
    UNIMM_CHECK(p->valid, "packer holds no batch");
    *out = p->desc.data();
    *n = static_cast<int32_t>(p->desc.size());
    return 0;
}

int unimm_packer_pack(unimm_packer_t* p, const unimm_flat_batch_t* fb, int threads) {
    UNIMM_CHECK(p != nullptr && fb != nullptr, "null argument");
    p->valid = false;
    const int S = p->S, R = p->R, F = p->F;
    const int U = fb->n_units, NB = fb->n_blocks;
    UNIMM_CHECK(U > 0 && NB > 0 && fb->blocks && fb->unit_block && fb->unit_row0 && fb->unit_rows, "packer: empty or incomplete flat batch");
    if (threads <= 0) threads = 4;
    const bool scores_only = fb->scores_only != 0, share_b0 = scores_only && fb->share_first_mask != 0;
    const int n_cls = scores_only ? 0 : 1;
    // ---- candidates per unit, argument checks
    std::vector<int> cand0(U + 1, 0);
    for (int u = 0; u < U; ++u) {
        const int b = fb->unit_block[u];
        UNIMM_CHECK(b >= 0 && b < NB && (u == 0 || b >= fb->unit_block[u - 1]), "packer: unit_block must be non-decreasing block indices");
        const unimm_image_block_t& blk = fb->blocks[b];
        UNIMM_CHECK(blk.input_ids && blk.token_type_ids && blk.position_ids && blk.masked_lm_labels && blk.image_feat && blk.image_loc &&
                        blk.image_mask, "packer: block with a NULL array");
        UNIMM_CHECK(fb->unit_rows[u] > 0 && fb->unit_row0[u] >= 0 && fb->unit_row0[u] + fb->unit_rows[u] <= blk.rows,
                    "packer: unit row range outside its block");
        cand0[u + 1] = cand0[u] + fb->unit_rows[u];
    }
    const int C = cand0[U];
    p->desc.assign(C, unimm_seq_desc_t{0, 0, 0, 0});
    p->cand_rep.assign(C, 0); p->cand_na.assign(C, 0); p->cand_nb.assign(C, 0);
    p->plan.assign(U, UnitPlan());
    p->unit_err.assign(U, std::string());

    // ---- pass 1 (parallel over units): descriptors, checks, row counts
    parallel_units(U, threads, [&](int u) {
        const unimm_image_block_t& blk = fb->blocks[fb->unit_block[u]];
        const int n = fb->unit_rows[u], r0 = fb->unit_row0[u], c0 = cand0[u];
        UnitPlan& pl = p->plan[u];
        std::string& err = p->unit_err[u];
        auto row = [&](const int64_t* a, int j) { return a + static_cast<size_t>(r0 + j) * S; };
        // descriptors: given, or derived from the position ids — the masked copy B repeats A's positions (data_utils.py:227), so
        // the first index whose position is not the index itself is L and its position is ctx
        int ctx_known = -1;
        for (int j = 0; j < n; ++j) {
            unimm_seq_desc_t d;
            if (blk.desc != nullptr) {
                d = blk.desc[r0 + j];
            } else {
                const int64_t* pos = row(blk.position_ids, j);
                int L = 1;
                while (L < S && pos[L] == L) ++L;
                d.mode = 0;
                if (L < S) { d.L = L; d.ctx = static_cast<int>(pos[L]); d.last_len = L - d.ctx; }
                else { d.L = S; d.ctx = -1; d.last_len = 0; }          // no masked copy inside S: ctx from a sibling below
            }
            if (d.ctx >= 0 && ctx_known < 0) ctx_known = d.ctx;
            p->desc[c0 + j] = d;
        }
        for (int j = 0; j < n; ++j) {
            unimm_seq_desc_t& d = p->desc[c0 + j];
            if (d.ctx < 0) {
                if (ctx_known < 0) { err = "cannot derive the context length of a unit whose candidates are all truncated before the masked copy: pass desc"; return; }
                d.ctx = ctx_known; d.last_len = d.L - d.ctx;
            }
            if (d.mode != 0) { err = "prefix sharing applies to generative-mode sequences only"; return; }
            if (d.ctx != p->desc[c0].ctx || d.ctx < 2) { err = "all candidates of a unit must share one context of at least one token"; return; }
            if (d.last_len < 1 || d.L != d.ctx + d.last_len) { err = "descriptor with L != ctx + last_len or an empty answer copy"; return; }
        }
        const int ctx = p->desc[c0].ctx;
        pl.ctx = ctx;
        pl.sh_len = std::min(ctx, S) - 1;
        if (fb->verify_shared && n > 1 && pl.sh_len > 0) {
            const size_t bytes = static_cast<size_t>(pl.sh_len) * sizeof(int64_t);
            for (int j = 1; j < n; ++j)
                if (std::memcmp(row(blk.input_ids, j) + 1, row(blk.input_ids, 0) + 1, bytes) != 0 ||
                    std::memcmp(row(blk.token_type_ids, j) + 1, row(blk.token_type_ids, 0) + 1, bytes) != 0 ||
                    std::memcmp(row(blk.position_ids, j) + 1, row(blk.position_ids, 0) + 1, bytes) != 0) {
                    err = "candidates of a unit differ in their context rows: cannot share the prefix";
                    return;
                }
        }
        // B_0 = [MASK] at position ctx seeing the context and itself: one row per unit when token / segment / position agree
        int b0 = (share_b0 && n > 1) ? 1 : 0;
        for (int j = 0; j < n && b0; ++j) {
            const unimm_seq_desc_t& d = p->desc[c0 + j];
            const int L0 = p->desc[c0].L;
            if (d.L >= S || row(blk.input_ids, j)[d.L] != row(blk.input_ids, 0)[L0] || row(blk.token_type_ids, j)[d.L] != row(blk.token_type_ids, 0)[L0] ||
                row(blk.position_ids, j)[d.L] != row(blk.position_ids, 0)[L0]) b0 = 0;
        }
        pl.b0_shared = b0;
        long long pairs = static_cast<long long>(pl.sh_len) * pl.sh_len + (b0 ? ctx : 0);
        for (int j = 0; j < n; ++j) {
            const unimm_seq_desc_t& d = p->desc[c0 + j];
            const int nb = std::max(0, std::min(d.last_len, S - d.L));
            const int na_av = std::max(0, std::min(d.last_len, S - d.ctx));
            const int na = scores_only ? std::min(na_av, std::max(nb - 1, 0)) : na_av;
            const int rep = n_cls + na + (nb - (nb > 0 ? b0 : 0));
            p->cand_na[c0 + j] = na; p->cand_nb[c0 + j] = nb; p->cand_rep[c0 + j] = rep;
            pl.own_rows += rep;
            pl.n_lm += nb;
            pl.n_lm_unique += nb - (nb > 0 ? b0 : 0);
            pl.max_rep = std::max(pl.max_rep, rep);
            // (query row, key) pairs: every own row sees the context; [CLS] all own rows, A_k k+1 keys, B_k k+1 keys (A_0..A_{k-1}, itself)
            pairs += static_cast<long long>(rep) * pl.sh_len + static_cast<long long>(n_cls) * rep;
            pairs += static_cast<long long>(na) * (na + 1) / 2;
            for (int k = b0; k < nb; ++k) pairs += k + 1;
        }
        pl.n_lm_unique += b0;
        pl.pairs = pairs;
    });
    for (int u = 0; u < U; ++u)
        if (!p->unit_err[u].empty()) { unimm::set_error("packer: unit " + std::to_string(u) + ": " + p->unit_err[u]); return 1; }

    // ---- prefix sums
    int n_shared = 0, n_lm = 0, n_ulm = 0, max_sh = 0, max_rep = 0, max_unit_rows = 0;
    long long pairs_ts = 0, sh_sum = 0;
    for (int u = 0; u < U; ++u) {
        UnitPlan& pl = p->plan[u];
        pl.sh_start = n_shared; n_shared += pl.sh_len;
        pl.cand0 = cand0[u];
        pl.lm0 = n_lm; n_lm += pl.n_lm;
        pl.ulm0 = n_ulm; n_ulm += pl.n_lm_unique;
        max_sh = std::max(max_sh, pl.sh_len);
        max_rep = std::max(max_rep, pl.max_rep);
        pairs_ts += pl.pairs;
        sh_sum += pl.sh_len;
    }
    int M = n_shared;
    for (int u = 0; u < U; ++u) {
        UnitPlan& pl = p->plan[u];
        pl.base = M;
        M += pl.b0_shared + pl.own_rows;
        max_unit_rows = std::max(max_unit_rows, pl.b0_shared + pl.own_rows);
    }
    max_rep = std::max(max_rep, 1);
    UNIMM_CHECK(max_sh <= 256, "packer: context longer than 256 rows");

    // ---- carve the arenas (every array 16-byte aligned)
    size_t io = 0;
    auto carve = [&](size_t n) { const size_t at = io; io += (n + 3) & ~size_t(3); return at; };
    const size_t o_ids = carve(M), o_seg = carve(M), o_pos = carve(M), o_iv = carve(static_cast<size_t>(M) * 4);
    const size_t o_jts = carve(static_cast<size_t>(2 * U) * 8), o_jt2i = carve(static_cast<size_t>(2 * U) * 8), o_ji2t = carve(static_cast<size_t>(U) * 8),
                 o_jimg = carve(static_cast<size_t>(U) * 8);
    const size_t o_lmr = carve(n_lm), o_lml = carve(n_lm), o_off = carve(static_cast<size_t>(C) + 1), o_cls = carve(C), o_img = carve(C);
    const size_t o_ur = carve(n_ulm), o_ui = carve(n_lm), o_uimg = carve(U);
    const size_t fneed = static_cast<size_t>(NB) * R * (F + 5 + 1) + 16;
    UNIMM_TRY(p->grow(reinterpret_cast<void**>(&p->ibuf), &p->icap, &p->ipinned, io + 16, sizeof(int32_t)));
    UNIMM_TRY(p->grow(reinterpret_cast<void**>(&p->fbuf), &p->fcap, &p->fpinned, fneed, sizeof(float)));
    int32_t* I = p->ibuf;
    int32_t *ids = I + o_ids, *seg = I + o_seg, *pos = I + o_pos, *iv = I + o_iv, *jts = I + o_jts, *jt2i = I + o_jt2i, *ji2t = I + o_ji2t,
            *jimg = I + o_jimg, *lmr = I + o_lmr, *lml = I + o_lml, *off = I + o_off, *cls = I + o_cls, *img = I + o_img, *ur = I + o_ur,
            *ui = I + o_ui, *uimg = I + o_uimg;
    float* feat = p->fbuf;
    float* loc = feat + static_cast<size_t>(NB) * R * F;
    float* msk = loc + static_cast<size_t>(NB) * R * 5;

    // ---- pass 2 (parallel over units, then over image blocks): fill
    std::atomic<int> bad_label{-1};
    parallel_units(U + NB, threads, [&](int w) {
        if (w >= U) {                                  // image blocks: one copy per IMAGE (val_lm.py:84-93 expands it x1000)
            const int b = w - U;
            const unimm_image_block_t& blk = fb->blocks[b];
            std::memcpy(feat + static_cast<size_t>(b) * R * F, blk.image_feat, sizeof(float) * R * F);
            std::memcpy(loc + static_cast<size_t>(b) * R * 5, blk.image_loc, sizeof(float) * R * 5);
            std::memcpy(msk + static_cast<size_t>(b) * R, blk.image_mask, sizeof(float) * R);
            return;
        }
        const int u = w;
        const UnitPlan& pl = p->plan[u];
        const int b = fb->unit_block[u];
        const unimm_image_block_t& blk = fb->blocks[b];
        const int n = fb->unit_rows[u], r0 = fb->unit_row0[u], c0 = pl.cand0;
        auto row = [&](const int64_t* a, int j) { return a + static_cast<size_t>(r0 + j) * S; };
        // context rows: dense columns 1 .. ctx-1 of candidate 0
        {
            const int64_t *t = row(blk.input_ids, 0), *s = row(blk.token_type_ids, 0), *q = row(blk.position_ids, 0);
            for (int i = 0; i < pl.sh_len; ++i) {
                const int r = pl.sh_start + i;
                ids[r] = static_cast<int32_t>(t[1 + i]); seg[r] = static_cast<int32_t>(s[1 + i]); pos[r] = static_cast<int32_t>(q[1 + i]);
                iv[4 * r] = 0; iv[4 * r + 1] = 0; iv[4 * r + 2] = -1; iv[4 * r + 3] = 0;
            }
        }
        const int q0 = pl.base, unit_rows = pl.b0_shared + pl.own_rows;
        if (pl.b0_shared) {
            const int L0 = p->desc[c0].L;
            ids[q0] = static_cast<int32_t>(row(blk.input_ids, 0)[L0]); seg[q0] = static_cast<int32_t>(row(blk.token_type_ids, 0)[L0]);
            pos[q0] = static_cast<int32_t>(row(blk.position_ids, 0)[L0]);
            iv[4 * q0] = q0; iv[4 * q0 + 1] = q0; iv[4 * q0 + 2] = q0; iv[4 * q0 + 3] = 0;     // no own-candidate keys besides itself
        }
        int r = q0 + pl.b0_shared, lm = pl.lm0, ulm = pl.ulm0;
        const int ub0 = ulm;                         // index of the shared B_0 row among the distinct labelled rows
        if (pl.b0_shared) ur[ulm++] = q0;
        for (int j = 0; j < n; ++j) {
            const unimm_seq_desc_t& d = p->desc[c0 + j];
            const int na = p->cand_na[c0 + j], nb = p->cand_nb[c0 + j], rep = p->cand_rep[c0 + j];
            const int b0 = nb > 0 ? pl.b0_shared : 0;
            const int64_t *t = row(blk.input_ids, j), *s = row(blk.token_type_ids, j), *q = row(blk.position_ids, j), *lab = row(blk.masked_lm_labels, j);
            const int first = r;
            auto put = [&](int dst, int col, int lo, int hi, int self) {
                ids[dst] = static_cast<int32_t>(t[col]); seg[dst] = static_cast<int32_t>(s[col]); pos[dst] = static_cast<int32_t>(q[col]);
                iv[4 * dst] = lo; iv[4 * dst + 1] = hi; iv[4 * dst + 2] = self; iv[4 * dst + 3] = 0;
            };
            cls[c0 + j] = n_cls ? first : -1;
            img[c0 + j] = u * R;
            off[c0 + j] = lm;
            if (n_cls) { put(r, 0, first, first + rep, -1); ++r; }
            const int a0 = r;
            for (int k = 0; k < na; ++k, ++r) put(r, d.ctx + k, a0, a0 + k + 1, -1);
            for (int k = 0; k < nb; ++k) {
                const int64_t y = lab[d.L + k];
                if (y < 0) bad_label.store(u);
                lml[lm] = static_cast<int32_t>(y);
                if (k == 0 && b0) { lmr[lm] = q0; ui[lm] = ub0; ++lm; continue; }
                put(r, d.L + k, a0, a0 + k, r);
                lmr[lm] = r; ui[lm] = ulm; ur[ulm++] = r;
                ++lm; ++r;
            }
        }
        const int s0 = pl.sh_start;
        auto job = [](int32_t* j, int a, int b_, int c, int d, int e, int f) { j[0] = a; j[1] = b_; j[2] = c; j[3] = d; j[4] = e; j[5] = f; j[6] = 0; j[7] = 0; };
        job(jts + 8 * u, s0, pl.sh_len, s0, pl.sh_len, 0, -1);                  // context rows attend their context
        job(jts + 8 * (U + u), q0, unit_rows, s0, pl.sh_len, 1, -1);            // candidate rows: context + own-candidate window
        job(jt2i + 8 * (2 * u), s0, pl.sh_len, u * R, R, 0, b);                 // text rows over the unit's image rows (mask row = image)
        job(jt2i + 8 * (2 * u + 1), q0, unit_rows, u * R, R, 0, b);
        job(ji2t + 8 * u, u * R, R, s0, pl.sh_len, 0, -1);                      // image rows over the unit's context rows
        job(jimg + 8 * u, u * R, R, u * R, R, 0, b);
        uimg[u] = b;
    });
    off[C] = n_lm;
    if (bad_label.load() >= 0) { unimm::set_error("packer: unit " + std::to_string(bad_label.load()) + ": a masked-copy position carries no label"); return 1; }

    unimm_packed_batch_t& o = p->batch;
    std::memset(&o, 0, sizeof(o));
    o.n_units = U; o.n_cands = C; o.n_text_rows = M;
    o.d_input_ids = ids; o.d_token_type_ids = seg; o.d_position_ids = pos; o.d_row_iv = iv;
    o.d_image_feat = feat; o.d_image_loc = loc; o.d_image_mask = msk;
    o.d_jobs_text_self = jts; o.n_jobs_text_self = 2 * U; o.n_jobs_text_ctx = U;
    o.max_q_text_self = std::max(std::max(max_unit_rows, 1), max_sh);
    o.cand_halo = max_rep - 1;
    o.d_jobs_t2i = jt2i; o.n_jobs_t2i = 2 * U; o.max_q_t2i = o.max_q_text_self;
    o.d_jobs_i2t = ji2t; o.n_jobs_i2t = U;
    o.d_jobs_img_self = jimg; o.n_jobs_img_self = U;
    o.kv_cap_text = roundup(std::max(max_sh, 1), 64);
    o.win_cap = roundup(Q_TILE + 2 * (max_rep - 1), 64);
    o.d_lm_rows = lmr; o.d_lm_labels = lml; o.n_lm_rows = n_lm;
    o.d_cand_lm_off = off; o.d_cand_cls_row = cls; o.d_cand_img_row = img;
    o.pairs_text_self = static_cast<double>(pairs_ts);
    o.pairs_i2t = static_cast<double>(R) * static_cast<double>(sh_sum);
    o.n_shared_rows = n_shared;
    o.no_cls_rows = scores_only ? 1 : 0;
    o.d_lm_urows = ur; o.d_lm_uidx = ui; o.n_lm_unique = n_ulm < n_lm ? n_ulm : 0;
    o.d_unit_image = uimg; o.n_images = NB;
    p->valid = true;
    return 0;
}

}  // extern "C"
