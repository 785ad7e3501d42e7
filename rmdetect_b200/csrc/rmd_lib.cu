// rmd_lib.cu — the C ABI declared in include/rmdetect_b200.h: context, device memory, launches.
//
// One context = one GPU, one compute stream, one copy stream.  Calls on a context are serialised
// by the caller (the reference itself is single-threaded and not re-entrant, lib/node.py:178-181).
// Nothing here computes on the host: without a CUDA device rmd_create() fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "scan_kernels.cuh"
#include "post_kernels.cuh"
#include "calib_kernels.cuh"
#include "synth_kernels.cuh"
#include "table_kernels.cuh"

static thread_local std::string g_error;

// The model tables live in __constant__ memory, which is per device, not per context.  Every context keeps host
// copies (rmd_ctx::sym) and `bind_tables` re-uploads them when another context (or a newer generation of this
// one) owns the device's symbols; calls on contexts of one device are serialised by the device's mutex.
#define RMD_MAX_DEVICES 64
static std::recursive_mutex g_dev_mutex[RMD_MAX_DEVICES];
static rmd_ctx *g_sym_owner[RMD_MAX_DEVICES];
static unsigned long long g_sym_gen[RMD_MAX_DEVICES];
static unsigned long long g_next_gen = 1;

// Tuning experiments (team count, CTAs per SM, kernel variants) are compiled in only with -DRMD_TUNING: the
// release library reads no environment variable, so a process that inherits RMD_* cannot change what runs.
#ifdef RMD_TUNING
static const char *tune_env(const char *name) { return getenv(name); }
#else
static const char *tune_env(const char *) { return nullptr; }
#endif

struct SymbolTables {              // host copies of the __constant__ tables other than c_models (rmd_ctx::hmodels)
    TrieNode trie[TRIE_CAP];
    uint16_t prel[32];
    int n_prel;
    unsigned char groups[RMD_MAX_MODELS + 1];
    int n_groups;
    ScreenBuild build[SCREEN_CAP];
    ScreenDesc desc[SCREEN_CAP];
    CandNode cand[TRIE_CAP];
    uint32_t ancw[TRIE_CAP];
    unsigned char cgrp[RMD_MAX_MODELS + 1][4];
};

struct DBuf {                      // grow-only device buffer
    void *p = nullptr;
    size_t cap = 0;
};

// device-resident hit table (rmd_table_*, rmd_table.cuh): rows plus the scratch of its sorts and selections
struct HitTable {
    DBuf rows, rows_alt, keys, keys_alt, idx, idx_alt, flag, off, aux32, aux64, hist, hist_off, totals, ranks, bmax;
    DBuf col_off, cols, grp, cl_info, cl_of_row, members, sums, counts;
    long long n = 0, n_clusters = 0;
    uint32_t next_id = 0;
    int min_key = 0x7fffffff, max_key = -1;        // sequence ids present
};

struct rmd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t stream2 = nullptr;     // streamed scans without expansion launches: chunks alternate between two compute streams, so
                                        // the next chunk's CTAs take the SMs the previous launch's tail leaves idle
    cudaEvent_t ev[16] = {};
    std::string error;
    // models
    int n_models = 0;
    std::vector<CModel> hmodels;
    SymbolTables sym{};
    unsigned long long sym_gen = 0;     // bumped whenever hmodels / sym change; compared with the device's g_sym_gen
    std::vector<void *> model_allocs;
    std::vector<void *> ev_allocs[RMD_MAX_MODELS];
    int max_chain = 0;
    int n_xnodes = 0, n_xcpt = 0, sm_count = 0;
    bool motif_ok = true;
    bool any_brute = false;             // some model gives no anchor for candidates: scan_fast_kernel<true>
    DBuf d_xnodes, d_xcpt, d_next;
    int teams = 8, ctas_per_sm = 1;      // shape of the persistent scan kernel (rmd_set_models)
    size_t scan_smem = 0;
    int teams_big = 0;                   // the same kernel staged for windows of 161..256 letters (FastSmemT<256>): teams per CTA, 0: not available
    size_t scan_smem_big = 0;
    bool tables_in_smem = true;
    // prefix screens (device_common.cuh: ScreenDesc): per model [screen0[m], screen0[m + 1]), words of one sequence's tables
    int n_screens = 0;
    unsigned char screen0[RMD_MAX_MODELS + 1] = {0};
    long long screen_words = 0;
    bool screens_dirty = true;          // GC tables, cutoff or models changed since the tables were built
    long long screen_min_windows = 1024; // sequences with fewer windows are scanned without screens; < 0: no screens at all
    DBuf d_screen_bits, d_screen_off, d_screen_seqs;
    long long strand_total = 0;         // rmd_synth_strand: letters of the whole synthetic stream (0: unbounded forward stream)
    bool strand_reverse = false;
    bool local_stack = false;           // gap trees branch deeper than STK_SLOTS: the engine's save slots live in local memory
    // resident job
    bool have_job = false;
    DevJob job{};
    long long n_letters = 0, n_bpp = 0;
    int max_wlen = 0, min_wlen = 0;
    bool any_other = false;
    std::vector<long long> h_seq_off, h_win_base;
    DBuf d_seq_off, d_win_base, d_codes8, d_codes2, d_gc_log, d_gc_class, d_bpp_off, d_bpp_i, d_bpp_j, d_bpp_p;
    DBuf d_flag, d_counts, d_bpp_cnt, d_bpp_ij, d_bpp_q;      // d_bpp_ij / d_bpp_q: compact transport staging
    // scan outputs
    DBuf d_other;                       // work list of the OTHER launch: window numbers, then their count
    DBuf d_raw, d_counters, d_gate_pass, d_group_hits, d_group_off, d_scan_sums, d_final, d_keep, d_keep_off, d_compact;
    rmd_hit *h_hits = nullptr;          // pinned; holds rmd_hit32 records when compact_hits
    size_t h_hits_cap = 0;
    bool compact_hits = false;          // rmd_set_hit_format
    DBuf d_pack;
    uint32_t *h_gate = nullptr, *h_group = nullptr;
    double *h_calcls = nullptr;         // pinned: a session's class logs on their way to the device (CAL_MAX_CLASS x 4)
    double *h_calhist = nullptr;        // pinned: rmd_calibrate's download of a session's histogram
    size_t h_calhist_n = 0;
    unsigned int *h_ready = nullptr;    // pinned: progress marks of a streamed scan (windows whose lists have been copied)
    size_t h_gate_cap = 0;
    long long raw_cap = 0;
    int stream_fallbacks = 0;           // streamed scans repeated as plain launches because the copies could not overlap (see scan_core)
    bool return_dropped = false;
    bool filter_score = false;          // rmd_set_min_score: drop hits with score <= min_score before the download
    double min_score = 0.0;
    // eval / calibrate scratch
    DBuf d_e0, d_e1, d_e2, d_e3, d_e4, d_e5, d_e6, d_e7, d_caltab, d_calcells, d_calcls, d_calthr, d_calhist, d_calcnt, d_calok;
    int cal_thr_bins = 0;                          // rounding thresholds resident in d_calthr for this many buckets
    size_t cal_smem = 0;                           // calibrate_kernel's count histogram in shared memory (0: in global memory)
    // calibration session (rmd_calib_open .. rmd_calib_read): buffers of its own (d_cal*), so that rmd_eval calls in between leave it alone
    CalibArgs cal{};
    bool cal_open = false, cal_timed = false;
    // the last scan's kept hits, still on the device (rmd_table_append_scan)
    const rmd_hit *last_hits = nullptr;
    long long last_n = 0;
    HitTable table;
};

static int fail(rmd_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    if (ctx) ctx->error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, RMD_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static int reserve(rmd_ctx *ctx, DBuf &b, size_t bytes) {
    if (bytes <= b.cap) return RMD_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) return fail(ctx, RMD_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    b.cap = want;
    return RMD_OK;
}
#define RESERVE(buf, bytes) do { int r_ = reserve(ctx, buf, bytes); if (r_) return r_; } while (0)

extern "C" const char *rmd_version(void) { return "rmdetect_b200 0.1.0 (sm_100a)"; }

extern "C" const char *rmd_last_error(rmd_ctx *ctx) { return ctx ? ctx->error.c_str() : g_error.c_str(); }

extern "C" int rmd_create(int device, rmd_ctx **out) {
    rmd_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, RMD_E_INVALID, "rmd_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, RMD_E_CUDA, "no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count || device >= RMD_MAX_DEVICES) return fail(nullptr, RMD_E_INVALID, "device %d out of range (%d present)", device, count);
    ctx = new rmd_ctx();
    ctx->device = device;
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    for (auto &ev : ctx->ev) CK(cudaEventCreate(&ev));
    *out = ctx;
    return RMD_OK;
}

static void free_models(rmd_ctx *ctx) {
    for (void *p : ctx->model_allocs) cudaFree(p);
    ctx->model_allocs.clear();
    for (auto &v : ctx->ev_allocs) { for (void *p : v) cudaFree(p); v.clear(); }
    ctx->hmodels.clear();
    ctx->n_models = 0;
}

extern "C" void rmd_destroy(rmd_ctx *ctx) {
    if (!ctx) return;
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (g_sym_owner[ctx->device] == ctx) g_sym_owner[ctx->device] = nullptr;   // its table pointers die with it
    free_models(ctx);
    DBuf *bufs[] = {&ctx->d_seq_off, &ctx->d_win_base, &ctx->d_codes8, &ctx->d_codes2, &ctx->d_gc_log, &ctx->d_gc_class,
                    &ctx->d_bpp_off, &ctx->d_bpp_i, &ctx->d_bpp_j, &ctx->d_bpp_p, &ctx->d_flag, &ctx->d_counts, &ctx->d_bpp_cnt, &ctx->d_bpp_ij, &ctx->d_bpp_q,
                    &ctx->d_raw, &ctx->d_counters, &ctx->d_gate_pass, &ctx->d_group_hits, &ctx->d_group_off, &ctx->d_scan_sums,
                    &ctx->d_final, &ctx->d_keep, &ctx->d_keep_off, &ctx->d_compact, &ctx->d_pack,
                    &ctx->d_xnodes, &ctx->d_xcpt, &ctx->d_next, &ctx->d_other, &ctx->d_e0, &ctx->d_e1, &ctx->d_e2, &ctx->d_e3, &ctx->d_e4, &ctx->d_e5, &ctx->d_e6, &ctx->d_e7, &ctx->d_caltab, &ctx->d_calcells, &ctx->d_calcls, &ctx->d_calthr, &ctx->d_calhist, &ctx->d_calcnt, &ctx->d_calok};
    for (DBuf *b : bufs) if (b->p) cudaFree(b->p);
    HitTable &T = ctx->table;
    DBuf *tbufs[] = {&T.rows, &T.rows_alt, &T.keys, &T.keys_alt, &T.idx, &T.idx_alt, &T.flag, &T.off, &T.aux32, &T.aux64, &T.hist, &T.hist_off, &T.totals,
                     &T.ranks, &T.bmax, &T.col_off, &T.cols, &T.grp, &T.cl_info, &T.cl_of_row, &T.members, &T.sums, &T.counts};
    for (DBuf *b : tbufs) if (b->p) cudaFree(b->p);
    if (ctx->h_hits) cudaFreeHost(ctx->h_hits);
    if (ctx->h_gate) cudaFreeHost(ctx->h_gate);
    if (ctx->h_group) cudaFreeHost(ctx->h_group);
    if (ctx->h_ready) cudaFreeHost(ctx->h_ready);
    if (ctx->h_calhist) cudaFreeHost(ctx->h_calhist);
    if (ctx->h_calcls) cudaFreeHost(ctx->h_calcls);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

extern "C" void *rmd_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void rmd_host_free(void *p) { if (p) cudaFreeHost(p); }

template <class T>
static int to_device(rmd_ctx *ctx, const T *src, size_t n, T **dst, std::vector<void *> &owner) {
    void *p = nullptr;
    CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    owner.push_back(p);
    if (n) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (T *)p;
    return RMD_OK;
}

extern "C" int rmd_set_models(rmd_ctx *ctx, const rmd_model_desc *models, int n_models) {
    if (!ctx || !models || n_models < 1) return fail(ctx, RMD_E_INVALID, "rmd_set_models: bad arguments");
    if (n_models > RMD_MAX_MODELS) return fail(ctx, RMD_E_LIMIT, "at most %d models per context", RMD_MAX_MODELS);
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    // Everything is validated and staged into temporaries first: a descriptor that is refused leaves the context
    // (its models, its resident job) exactly as it was.
    struct Pending { std::vector<void *> allocs; bool committed = false; ~Pending() { if (!committed) for (void *p : allocs) cudaFree(p); } } pending;
    std::vector<CModel> hmodels((size_t)n_models, CModel{});
    std::unique_ptr<SymbolTables> symp(new SymbolTables());
    SymbolTables &sym = *symp;
    memset(&sym, 0, sizeof sym);
    int max_chain = 0;
    int trie_total = 0;
    TrieNode *const htrie = sym.trie;
    std::vector<XNode> hx;
    bool motif_ok = true;
    uint16_t *const hprel = sym.prel;
    int n_prel = 0;
    bool prel_overflow = false;
    int stack_slots = 1;
    int n_screens = 0;
    long long screen_words = 0;
    unsigned char screen0[RMD_MAX_MODELS + 1] = {0};
    ScreenBuild *const hbuild = sym.build;
    ScreenDesc *const hdesc = sym.desc;
    std::vector<double> hcpt;
    for (int m = 0; m < n_models; m++) {
        const rmd_model_desc &d = models[m];
        CModel &c = hmodels[m];
        if (d.n_gate_pairs > RMD_MAX_GATE_PAIRS || d.n_gate_cfg > RMD_MAX_GATE_CFGS || d.n_first > RMD_MAX_FIRST_PAIRS)
            return fail(ctx, RMD_E_LIMIT, "model %d: gate program too large (%d pairs, %d configurations, %d first pairs)",
                        m, d.n_gate_pairs, d.n_gate_cfg, d.n_first);
        if (d.n_tree < 1 || d.n_tree > 65535 || d.chain_len[0] < 1 || d.chain_len[1] < 1 || d.chain_len[0] > 100 || d.chain_len[1] > 100)
            return fail(ctx, RMD_E_LIMIT, "model %d: tree size %d / chain lengths %d,%d out of range", m, d.n_tree, d.chain_len[0], d.chain_len[1]);
        if (!d.tree || !d.leaf_dist || !d.cpt || !d.gate_cfg || (d.n_gate_pairs > 0 && !d.gate_pairs) || !d.leaf_offsets)
            return fail(ctx, RMD_E_INVALID, "model %d: NULL table", m);
        if (d.n_gate_cfg < 0 || d.n_gate_pairs < 0 || d.n_leaves < 1 || d.n_cpt_rows < 1)
            return fail(ctx, RMD_E_INVALID, "model %d: negative or empty table size", m);
        if (d.gate_cfg[0] != 0 || d.gate_cfg[d.n_gate_cfg] != d.n_gate_pairs)
            return fail(ctx, RMD_E_INVALID, "model %d: gate_cfg does not span the %d gate pairs", m, d.n_gate_pairs);
        for (int k = 0; k < d.n_gate_cfg; k++)
            if (d.gate_cfg[k + 1] < d.gate_cfg[k]) return fail(ctx, RMD_E_INVALID, "model %d: gate_cfg is not monotone", m);
        c.chain_len[0] = d.chain_len[0]; c.chain_len[1] = d.chain_len[1];
        max_chain = std::max(max_chain, std::max(d.chain_len[0], d.chain_len[1]));
        c.symmetric = d.symmetric; c.n_tree = d.n_tree; c.n_cfg = d.n_gate_cfg;
        c.brute = d.brute_candidates; c.n_leaves = d.n_leaves;
        for (int k = 0; k <= d.n_gate_cfg; k++) c.cfg_begin[k] = (unsigned short)d.gate_cfg[k];
        for (int k = 0; k < d.n_gate_pairs; k++) {
            c.pairs[k] = d.gate_pairs[k];
            if ((c.pairs[k].c1 | c.pairs[k].c2) & ~1) return fail(ctx, RMD_E_INVALID, "model %d: gate pair names chain > 1", m);
        }
        c.n_first = d.n_first;
        {   // prefix trie over the configurations' pair sequences, emitted in DFS preorder
            struct Tmp { int ro, co, parent, mult; std::vector<int> kids; };
            std::vector<Tmp> tmp;
            std::vector<int> roots;
            for (int g = 0; g < d.n_gate_cfg; g++) {
                if (d.gate_cfg[g + 1] <= d.gate_cfg[g]) return fail(ctx, RMD_E_INVALID, "model %d: empty gate configuration", m);
                int parent = -1;
                for (int k = d.gate_cfg[g]; k < d.gate_cfg[g + 1]; k++) {
                    const rmd_gate_pair q = d.gate_pairs[k];
                    int ro, co;
                    if (q.c1 == q.c2) { c.brute = 1; ro = co = 0; }          // same-chain pair: no anchor for candidates
                    else if (q.c1 == 0) { ro = q.o1; co = q.o2; }
                    else { ro = q.o2; co = q.o1; }
                    std::vector<int> &sibs = parent < 0 ? roots : tmp[parent].kids;
                    int node = -1;
                    for (int t : sibs) if (tmp[t].ro == ro && tmp[t].co == co) { node = t; break; }
                    if (node < 0) {
                        node = (int)tmp.size();
                        tmp.push_back(Tmp{ro, co, parent, 0, {}});
                        (parent < 0 ? roots : tmp[parent].kids).push_back(node);
                    }
                    tmp[node].mult++;
                    parent = node;
                }
            }
            if (trie_total + (int)tmp.size() > TRIE_CAP)
                return fail(ctx, RMD_E_LIMIT, "gate programs of the model set too large for shared memory (%d trie nodes, limit %d)", trie_total + (int)tmp.size(), TRIE_CAP);
            c.trie0 = trie_total; c.n_trie = (int)tmp.size();
            std::vector<int> pos(tmp.size(), -1);
            std::vector<std::pair<int, int>> stack;               // (node, next child)
            for (int rt : roots) {
                stack.push_back({rt, 0});
                while (!stack.empty()) {
                    auto &top = stack.back();
                    const int nd = top.first;
                    if (top.second == 0) {
                        pos[nd] = trie_total++;
                        TrieNode &tn = htrie[pos[nd]];
                        if (tmp[nd].mult > 255) return fail(ctx, RMD_E_LIMIT, "model %d: more than 255 configurations share a pair", m);
                        tn.ro = (int8_t)tmp[nd].ro; tn.co = (int8_t)tmp[nd].co; tn.mult = (uint8_t)tmp[nd].mult; tn.model = (uint8_t)m;
                        tn.parent = (int8_t)(tmp[nd].parent < 0 ? -1 : pos[tmp[nd].parent]);
                        tn.pbit = 31;
                        if (tmp[nd].parent >= 0) {             // parent pair relative to this one
                            const int dr = tmp[tmp[nd].parent].ro - tmp[nd].ro, dc = tmp[tmp[nd].parent].co - tmp[nd].co;
                            const uint16_t key = (uint16_t)((dr & 0xff) | ((dc & 0xff) << 8));
                            int k = 0;
                            while (k < n_prel && hprel[k] != key) k++;
                            if (k == n_prel) {
                                if (n_prel == 31) { prel_overflow = true; k = 0; }
                                else hprel[n_prel++] = key;
                            }
                            tn.pbit = (uint8_t)k;
                        }
                    }
                    if (top.second < (int)tmp[nd].kids.size()) { const int kid = tmp[nd].kids[top.second++]; stack.push_back({kid, 0}); }
                    else { htrie[pos[nd]].skip = (unsigned short)trie_total; stack.pop_back(); }
                }
            }
        }
        int r;
        rmd_tree_node *tree = nullptr; int32_t *ld = nullptr; double *cpt = nullptr; int8_t *lo = nullptr;
        for (int t = 0; t < d.n_tree; t++)
            if (d.tree[t].npar > 2 || d.tree[t].bdepth > RMD_MAX_BRANCH_DEPTH || d.tree[t].leaf >= d.n_leaves ||
                (int)d.tree[t].cpt_row + (d.tree[t].npar == 2 ? RMD_NSYM * RMD_NSYM : d.tree[t].npar == 1 ? RMD_NSYM : 1) > d.n_cpt_rows ||
                (d.tree[t].chain & ~1) || d.tree[t].oft >= d.chain_len[d.tree[t].chain & 1])
                return fail(ctx, RMD_E_LIMIT, "model %d: tree node %d out of range", m, t);
        // level of every tree node (depth along ORDER), kept in the node's spare byte for the screen tabulation
        std::vector<rmd_tree_node> ltree(d.tree, d.tree + d.n_tree);
        std::vector<int> level((size_t)d.n_tree, 0);
        {
            std::vector<std::pair<int, int>> open;                  // (end of subtree, level) of the ancestors
            for (int t = 0; t < d.n_tree; t++) {
                while (!open.empty() && open.back().first <= t) open.pop_back();
                level[t] = open.empty() ? 0 : open.back().second + 1;
                if (d.tree[t].skip <= t || d.tree[t].skip > d.n_tree) return fail(ctx, RMD_E_INVALID, "model %d: tree node %d has a bad skip link", m, t);
                open.push_back({d.tree[t].skip, level[t]});
                ltree[t].pad_ = (uint8_t)std::min(level[t], 255);
            }
        }
        if ((r = to_device(ctx, ltree.data(), (size_t)d.n_tree, &tree, pending.allocs))) return r;
        screen0[m] = (unsigned char)n_screens;
        if (motif_ok && !tune_env("RMD_NO_SCREEN")) {
            // one screen per top-level subtree, all of the same depth: the deepest level whose letters fit
            // SCREEN_MAX_LETTERS (small trees: 6, their tables stay in L1) in at most SCREEN_MAX_RUNS runs of
            // neighbouring offsets
            const int max_letters = d.n_tree > 100 ? SCREEN_MAX_LETTERS : 6;
            std::vector<std::pair<int, int>> roots;
            for (int t = 0; t < d.n_tree; t = d.tree[t].skip) roots.push_back({t, (int)d.tree[t].skip});
            int deepest = 0;
            for (int t = 0; t < d.n_tree; t++) deepest = std::max(deepest, level[t]);
            auto letters_of = [&](int t0, int t1, int L) {
                std::vector<int> pos;
                auto add = [&](int chain, int oft) { if (oft >= 0) { const int k = 2 * oft + 32 * (chain ? 1 : 0); if (std::find(pos.begin(), pos.end(), k) == pos.end()) pos.push_back(k); } };
                for (int t = t0; t < t1; t++) {
                    if (level[t] > L) continue;
                    add(d.tree[t].chain, d.tree[t].oft);
                    for (int k = 0; k < d.tree[t].npar; k++) add(d.tree[t].par_chain[k], d.tree[t].par_oft[k]);
                }
                std::sort(pos.begin(), pos.end());
                return pos;
            };
            auto runs_of = [](const std::vector<int> &pos) { int n = 0; for (size_t k = 0; k < pos.size(); k++) if (k == 0 || pos[k] != pos[k - 1] + 2 || pos[k] == 32) n++; return n; };
            int best = -1;
            for (int L = 0; L <= std::min(deepest, 254); L++) {
                bool fits = true;
                for (auto &rt : roots) { auto pos = letters_of(rt.first, rt.second, L); if ((int)pos.size() > max_letters || runs_of(pos) > SCREEN_MAX_RUNS) fits = false; }
                if (!fits) break;
                best = L;
            }
            if (best >= 1 && (int)roots.size() <= SCREEN_MAX_PER_MODEL) {
                for (auto &rt : roots) {
                    const std::vector<int> pos = letters_of(rt.first, rt.second, best);
                    ScreenBuild &B = hbuild[n_screens];
                    ScreenDesc &D = hdesc[n_screens];
                    memset(&B, 0, sizeof B); memset(&D, 0, sizeof D);
                    B.model = m; B.t0 = rt.first; B.t1 = rt.second; B.level = best; B.npos = (int)pos.size();
                    B.word_off = D.word_off = (uint32_t)screen_words;
                    int run = -1;
                    for (size_t k = 0; k < pos.size(); k++) {
                        B.pos[k] = (uint8_t)pos[k];
                        if (k == 0 || pos[k] != pos[k - 1] + 2 || pos[k] == 32) { run++; D.shift[run] = (uint8_t)pos[k]; D.dst[run] = (uint8_t)(2 * k); D.mask[run] = 0; }
                        D.mask[run] = (D.mask[run] << 2) | 3u;
                    }
                    screen_words += std::max<long long>(1, (1LL << (2 * pos.size())) / 32);
                    n_screens++;
                }
            }
        }
        if ((r = to_device(ctx, d.leaf_dist, (size_t)d.n_leaves * 2, &ld, pending.allocs))) return r;
        if ((r = to_device(ctx, d.cpt, (size_t)d.n_cpt_rows * RMD_NSYM, &cpt, pending.allocs))) return r;
        if ((r = to_device(ctx, d.leaf_offsets, (size_t)d.n_leaves * (d.chain_len[0] + d.chain_len[1]), &lo, pending.allocs))) return r;
        {   // the scoring engine's view: pre-decoded nodes, absolute indices, all models concatenated
            const size_t x0 = hx.size(), cpt0 = hcpt.size();
            if (x0 + (size_t)d.n_tree >= XNODE_END) return fail(ctx, RMD_E_LIMIT, "search trees of the model set exceed %u nodes", XNODE_END - 1);
            if (d.n_leaves > 65535) return fail(ctx, RMD_E_LIMIT, "model %d: too many gap configurations", m);
            c.xnode0 = (int)x0;
            // save slots of the scoring engine: depths are renumbered from the model's first saving node, so the shipped
            // models need three slots.  A node that restores at renumbered depth 0 restarts from zero: it is a child of
            // the implicit root (the flattener gives those depth 1 and no saving node above them).
            int dmin = RMD_MAX_BRANCH_DEPTH + 1, dmax = -1;
            for (int t = 0; t < d.n_tree; t++)
                if (d.tree[t].flags & RMD_F_SAVE) { dmin = std::min<int>(dmin, d.tree[t].bdepth); dmax = std::max<int>(dmax, d.tree[t].bdepth); }
            if (dmax < 0) dmin = 1;                                // no saving node: only root children may restore
            for (int t = 0; t < d.n_tree; t++)
                if ((d.tree[t].flags & RMD_F_RESTORE) && d.tree[t].bdepth <= dmin && d.tree[t].bdepth != 1)
                    return fail(ctx, RMD_E_INVALID, "model %d: tree node %d restores a sum that was never saved", m, t);
            stack_slots = std::max(stack_slots, dmax - dmin + 1);
            hcpt.insert(hcpt.end(), d.cpt, d.cpt + (size_t)d.n_cpt_rows * RMD_NSYM);
            for (int t = 0; t < d.n_tree; t++) {
                const rmd_tree_node &n = d.tree[t];
                XNode x;
                memset(&x, 0, sizeof x);
                const int stride[2] = {n.npar == 1 ? RMD_NSYM : n.npar == 2 ? RMD_NSYM * RMD_NSYM : 0,      // row = c0, or c0 * 6 + c1
                                       n.npar == 2 ? RMD_NSYM : 0};
                size_t base = cpt0 + (size_t)n.cpt_row * RMD_NSYM;
                uint32_t sel = 0;
                if (n.oft < 0) sel |= XSEL_OWN_GAP;
                else { if (n.oft > XNODE_MAX_OFFSET) motif_ok = false; sel |= (uint32_t)((2 * n.oft + 32 * (n.chain ? 1 : 0)) & 63); }
                uint16_t mm[2] = {0, 0};
                for (int k = 0; k < n.npar; k++) {
                    if (n.par_oft[k] < 0) base += (size_t)4 * stride[k];                 // a gap parent: its letter is the gap code
                    else {
                        if (n.par_oft[k] > XNODE_MAX_OFFSET) motif_ok = false;
                        sel |= (uint32_t)((2 * n.par_oft[k] + 32 * (n.par_chain[k] ? 1 : 0)) & 63) << (8 + 8 * k);
                        mm[k] = (uint16_t)stride[k];
                    }
                }
                x.sel = sel; x.cpt = (uint32_t)base; x.m0 = mm[0]; x.m1 = mm[1];
                x.skip = n.skip >= d.n_tree ? (uint16_t)XNODE_END : (uint16_t)(x0 + n.skip);
                x.next = t + 1 >= d.n_tree ? (uint16_t)XNODE_END : (uint16_t)(x0 + t + 1);
                x.leaf = n.leaf; x.model = (uint8_t)m;
                x.fb = (uint8_t)((n.flags & 0xf) | (std::max(n.bdepth - dmin, 0) << 4));
                if (n.flags & RMD_F_CHECK_DIST) {
                    const int dx = d.leaf_dist[2 * n.leaf], dy = d.leaf_dist[2 * n.leaf + 1];
                    if (dx > 32767 || dy > 32767 || dx < 0 || dy < 0) return fail(ctx, RMD_E_LIMIT, "model %d: separation out of range", m);
                    // dead when 0 <= d < dx or -dy < d <= 0, d = p1 - p0: one interval [lo, hi)
                    const int lo = dy > 0 ? -dy + 1 : (dx > 0 ? 0 : 1), hi = dx > 0 ? dx : (dy > 0 ? 1 : 0);
                    x.dlo = (int16_t)lo; x.dspan = (uint16_t)(hi > lo ? hi - lo : 0);
                }
                hx.push_back(x);
            }
        }
        c.tree = tree; c.leaf_dist = (const int2 *)ld; c.cpt = cpt; c.leaf_offsets = lo;
        c.ev_table = nullptr; c.ev_thr = nullptr; c.ev_nclass = 0; c.ev_nrows = 0;
    }
    {   // groups of models whose candidates are generated together: a gate trie of more than two nodes needs a
        // duplicate bitmap, and CB_SLOTS bitmaps are resident at a time
        unsigned char groups[RMD_MAX_MODELS + 1] = {0};
        int n_groups = 0, used = 0;
        for (int m = 0; m < n_models; m++) {
            CModel &c = hmodels[m];
            const bool bitmap = c.n_trie > 2;
            if (bitmap && used == CB_SLOTS) { groups[++n_groups] = (unsigned char)m; used = 0; }
            c.cb_slot = bitmap ? used++ : -1;
        }
        groups[++n_groups] = (unsigned char)n_models;
        memcpy(sym.groups, groups, sizeof groups);
        sym.n_groups = n_groups;
        // candidate order (device_common.cuh: CandNode): per group the root nodes, then the others, each in trie order
        int k = 0;
        for (int t = 0; t < trie_total; t++)
            sym.ancw[t] = (uint32_t)(htrie[t].ro & 0xff) | ((uint32_t)(htrie[t].co & 0xff) << 12) | ((uint32_t)(htrie[t].parent < 0 ? 0x7f : htrie[t].parent) << 24);
        for (int g = 0; g < n_groups; g++) {
            const int t0 = hmodels[groups[g]].trie0, t1 = hmodels[groups[g + 1] - 1].trie0 + hmodels[groups[g + 1] - 1].n_trie;
            sym.cgrp[g][0] = (unsigned char)k;
            // models without SYMMETRIC first (their odometer range is one interval test per node), the symmetric ones after
            // cgrp[g][1]; inside each part the root nodes (first pairs of configurations), then the others, each in trie order
            for (int pass = 0; pass < 4; pass++) {
                if (pass == 2) sym.cgrp[g][1] = (unsigned char)k;
                for (int t = t0; t < t1; t++) {
                    const TrieNode &tn = htrie[t];
                    const CModel &c = hmodels[tn.model];
                    if ((c.symmetric != 0) != (pass >= 2) || (tn.parent < 0) != ((pass & 1) == 0)) continue;
                    if (tn.ro < 0 || tn.co < 0 || tn.ro > 100 || tn.co > 100) return fail(ctx, RMD_E_LIMIT, "model %d: gate pair offset out of range", (int)tn.model);
                    const int gp = tn.parent < 0 ? 0x7f : (htrie[tn.parent].parent < 0 ? 0x7f : htrie[tn.parent].parent);
                    CandNode &cn = sym.cand[k++];
                    cn.x = (uint32_t)(tn.ro + 16) | ((uint32_t)(tn.co + 16) << 12) | ((uint32_t)(tn.pbit & 31) << 24) | (c.symmetric ? NODE_SYM : 0u);
                    cn.info = (uint32_t)t | ((uint32_t)gp << 7) | ((uint32_t)tn.model << 14) | ((uint32_t)(c.cb_slot + 1) << 18) |
                              ((c.cb_slot < 0 && t != c.trie0) ? 1u << 20 : 0u) | ((uint32_t)c.trie0 << 21);
                }
            }
            sym.cgrp[g][2] = (unsigned char)k;
        }
    }
    sym.n_prel = n_prel;
    bool any_brute = prel_overflow;               // too many distinct pair geometries for the candidate filter: exact path
    for (const CModel &c : hmodels) if (c.brute) any_brute = true;
    for (int m = n_models; m <= RMD_MAX_MODELS; m++) screen0[m] = (unsigned char)n_screens;
    int teams_new = 8, ctas_new = 1, sm_count = 0, teams_big_new = 0;
    size_t scan_smem = 0, scan_smem_big = 0;
    bool tables_in_smem = true;
    {   // shape of the persistent scan kernel: tables + one FastSmem per team within the SM's shared memory
        int smem_max = 0;
        CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
        CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        const size_t screens = sizeof(ScreenDesc) * (size_t)n_screens;          // always in shared memory
        const size_t tables = sizeof(XNode) * hx.size() + 16 * ((hcpt.size() + 1) / 2);
        int teams = tune_env("RMD_TEAMS") ? atoi(tune_env("RMD_TEAMS")) : SCAN_MAX_THREADS / SCAN_THREADS;
        ctas_new = tune_env("RMD_CTAS_PER_SM") ? std::max(1, atoi(tune_env("RMD_CTAS_PER_SM"))) : 1;
        teams = std::max(1, std::min(teams, SCAN_MAX_THREADS / SCAN_THREADS));
        tables_in_smem = tables + screens + SCAN_CTA_FIXED + 2 * sizeof(FastSmem) <= (size_t)smem_max / ctas_new;
        const size_t fixed = (tables_in_smem ? tables : 0) + screens + SCAN_CTA_FIXED;
        while (teams > 1 && fixed + teams * sizeof(FastSmem) > (size_t)smem_max / ctas_new) teams--;
        teams_new = teams;
        scan_smem = fixed + teams * sizeof(FastSmem);
        if (scan_smem > (size_t)smem_max) return fail(ctx, RMD_E_LIMIT, "scan kernel needs %zu bytes of shared memory, the device offers %d", scan_smem, smem_max);
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        if (tune_env("RMD_CARVEOUT")) {                // L1 / shared memory split of the main kernel, per cent of shared memory
            CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(tune_env("RMD_CARVEOUT"))));
            CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(tune_env("RMD_CARVEOUT"))));
        }
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false, false, false, RMD_BIG_WINDOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, false, false, true, RMD_BIG_WINDOW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        // windows of 161..256 letters: the same kernel with a larger bit matrix per team, as many teams as still fit
        teams_big_new = 0;
        if (tables_in_smem) {
            int tb = SCAN_MAX_THREADS / SCAN_THREADS;
            while (tb > 0 && fixed + tb * sizeof(FastSmemT<RMD_BIG_WINDOW>) > (size_t)smem_max / ctas_new) tb--;
            teams_big_new = tb;
            scan_smem_big = fixed + tb * sizeof(FastSmemT<RMD_BIG_WINDOW>);
        }
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        CK(cudaFuncSetAttribute(scan_fast_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        CK(cudaFuncSetAttribute(scan_fast_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        CK(cudaFuncSetAttribute(scan_fast_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    }
    // ---- commit: nothing below can be refused for the descriptors' sake ------------------------------------------
    CK(cudaDeviceSynchronize());
    RESERVE(ctx->d_xnodes, sizeof(XNode) * std::max<size_t>(hx.size(), 1));
    RESERVE(ctx->d_xcpt, sizeof(double) * std::max<size_t>(hcpt.size(), 1));
    CK(cudaMemcpy(ctx->d_xnodes.p, hx.data(), sizeof(XNode) * hx.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_xcpt.p, hcpt.data(), sizeof(double) * hcpt.size(), cudaMemcpyHostToDevice));
    ctx->have_job = false;                         // a resident job was staged for the old model count
    free_models(ctx);
    ctx->model_allocs.swap(pending.allocs);
    pending.committed = true;
    ctx->hmodels.swap(hmodels);
    ctx->sym = sym;
    ctx->sym_gen = g_next_gen++;                   // bind_tables() uploads the __constant__ copies before the next launch
    ctx->n_models = n_models;
    ctx->max_chain = max_chain;
    ctx->any_brute = any_brute;
    ctx->n_xnodes = (int)hx.size(); ctx->n_xcpt = (int)hcpt.size(); ctx->motif_ok = motif_ok;
    ctx->local_stack = stack_slots > STK_SLOTS || tune_env("RMD_LOCAL_STACK");
    memcpy(ctx->screen0, screen0, sizeof screen0);
    ctx->n_screens = n_screens; ctx->screen_words = screen_words; ctx->screens_dirty = true;
    ctx->teams = teams_new; ctx->ctas_per_sm = ctas_new; ctx->sm_count = sm_count;
    ctx->scan_smem = scan_smem; ctx->tables_in_smem = tables_in_smem;
    ctx->teams_big = teams_big_new; ctx->scan_smem_big = scan_smem_big;
    return RMD_OK;
}

// Make the device's __constant__ tables this context's (see g_sym_owner).  Called, under the device lock, by every
// entry point that launches a kernel reading c_models / c_trie / c_prel / c_groups / c_screen*.
static int bind_tables(rmd_ctx *ctx) {
    const int dev = ctx->device;
    if (g_sym_owner[dev] == ctx && g_sym_gen[dev] == ctx->sym_gen) return RMD_OK;
    if (ctx->n_models == 0) return fail(ctx, RMD_E_STATE, "no models: call rmd_set_models first");
    CK(cudaDeviceSynchronize());                    // kernels of the previous owner may still be reading the symbols
    g_sym_owner[dev] = nullptr;
    const SymbolTables &t = ctx->sym;
    CK(cudaMemcpyToSymbol(c_models, ctx->hmodels.data(), sizeof(CModel) * ctx->n_models));
    CK(cudaMemcpyToSymbol(c_trie, t.trie, sizeof t.trie));
    CK(cudaMemcpyToSymbol(c_prel, t.prel, sizeof t.prel));
    CK(cudaMemcpyToSymbol(c_n_prel, &t.n_prel, sizeof t.n_prel));
    CK(cudaMemcpyToSymbol(c_groups, t.groups, sizeof t.groups));
    CK(cudaMemcpyToSymbol(c_n_groups, &t.n_groups, sizeof t.n_groups));
    CK(cudaMemcpyToSymbol(c_screen_build, t.build, sizeof t.build));
    CK(cudaMemcpyToSymbol(c_screens, t.desc, sizeof t.desc));
    CK(cudaMemcpyToSymbol(c_cand, t.cand, sizeof t.cand));
    CK(cudaMemcpyToSymbol(c_ancw, t.ancw, sizeof t.ancw));
    CK(cudaMemcpyToSymbol(c_cgrp, t.cgrp, sizeof t.cgrp));
    g_sym_owner[dev] = ctx; g_sym_gen[dev] = ctx->sym_gen;
    return RMD_OK;
}

extern "C" int rmd_set_evalues(rmd_ctx *ctx, int model, const double *table, int n_class, int n_rows, const double *thresholds) {
    if (!ctx || model < 0 || model >= ctx->n_models) return fail(ctx, RMD_E_INVALID, "rmd_set_evalues: bad model index");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->sym_gen = g_next_gen++;                  // c_models changes: re-bound before the next launch
    for (void *p : ctx->ev_allocs[model]) cudaFree(p);
    ctx->ev_allocs[model].clear();
    CModel &c = ctx->hmodels[model];
    c.ev_table = nullptr; c.ev_thr = nullptr; c.ev_nclass = 0; c.ev_nrows = 0;
    if (n_class > 0 && n_rows > 0) {
        if (!table || !thresholds) return fail(ctx, RMD_E_INVALID, "rmd_set_evalues: NULL table");
        double *t = nullptr, *th = nullptr; int r;
        if ((r = to_device(ctx, table, (size_t)n_class * n_rows, &t, ctx->ev_allocs[model]))) return r;
        if ((r = to_device(ctx, thresholds, (size_t)n_rows, &th, ctx->ev_allocs[model]))) return r;
        c.ev_table = t; c.ev_thr = th; c.ev_nclass = n_class; c.ev_nrows = n_rows;
    }
    return RMD_OK;
}

// ---- window bookkeeping (lib/scanner.py:80-111) ----------------------------------------------
static long long windows_of(long long len, int win_len, int win_step) {
    if (win_len == 0 || win_step == 0) return 1;
    if (len <= win_len) return 1;
    return 1 + (len - win_len + win_step - 1) / win_step;
}

static int stage_layout(rmd_ctx *ctx, int n_seq, const int64_t *seq_off, int win_len, int win_step, long long expect_win) {
    if (n_seq < 1 || !seq_off) return fail(ctx, RMD_E_INVALID, "job needs at least one sequence");
    if (win_len < 0 || win_step < 0) return fail(ctx, RMD_E_INVALID, "negative window length or step");
    ctx->h_seq_off.assign(seq_off, seq_off + n_seq + 1);
    ctx->h_win_base.assign(n_seq + 1, 0);
    ctx->max_wlen = 0;
    ctx->min_wlen = 0x7fffffff;
    for (int s = 0; s < n_seq; s++) {
        long long len = seq_off[s + 1] - seq_off[s];
        if (len < 0) return fail(ctx, RMD_E_INVALID, "seq_off is not monotone");
        ctx->h_win_base[s + 1] = ctx->h_win_base[s] + windows_of(len, win_len, win_step);
        long long wl = (win_len == 0 || win_step == 0) ? len : std::min<long long>(len, win_len);
        if (wl > 65535) return fail(ctx, RMD_E_LIMIT, "window of %lld letters exceeds the 65535 limit", wl);
        ctx->max_wlen = std::max(ctx->max_wlen, (int)wl);
        // shortest window of the sequence: the whole sequence if it fits one window, else the tail window is full length
        ctx->min_wlen = std::min(ctx->min_wlen, (int)wl);
    }
    if (expect_win >= 0 && ctx->h_win_base[n_seq] != expect_win)
        return fail(ctx, RMD_E_INVALID, "n_win is %lld but the sequences yield %lld windows", expect_win, ctx->h_win_base[n_seq]);
    ctx->n_letters = seq_off[n_seq];
    RESERVE(ctx->d_seq_off, sizeof(long long) * (n_seq + 1));
    RESERVE(ctx->d_win_base, sizeof(long long) * (n_seq + 1));
    CK(cudaMemcpyAsync(ctx->d_seq_off.p, ctx->h_seq_off.data(), sizeof(long long) * (n_seq + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_win_base.p, ctx->h_win_base.data(), sizeof(long long) * (n_seq + 1), cudaMemcpyHostToDevice, ctx->stream));
    DevJob &J = ctx->job;
    J.n_seq = n_seq; J.n_models = ctx->n_models; J.win_len = win_len; J.win_step = win_step;
    J.n_win = ctx->h_win_base[n_seq];
    J.debug = tune_env("RMD_DEBUG") ? atoi(tune_env("RMD_DEBUG")) : 0;      // -DRMD_TUNING builds only
    J.seq_off = (const long long *)ctx->d_seq_off.p; J.win_base = (const long long *)ctx->d_win_base.p;
    J.xnodes = (const XNode *)ctx->d_xnodes.p; J.xcpt = (const double *)ctx->d_xcpt.p;
    J.n_xnodes = ctx->n_xnodes; J.n_xcpt = ctx->n_xcpt; J.motif_ok = ctx->motif_ok ? 1 : 0;
    return RMD_OK;
}

static int stage_gc(rmd_ctx *ctx, const double *gc_log, const int32_t *gc_class) {
    DevJob &J = ctx->job;
    RESERVE(ctx->d_gc_log, sizeof(double) * RMD_N_CODES * J.n_seq);
    RESERVE(ctx->d_gc_class, sizeof(int) * (size_t)J.n_seq * J.n_models);
    if (gc_log) CK(cudaMemcpyAsync(ctx->d_gc_log.p, gc_log, sizeof(double) * RMD_N_CODES * J.n_seq, cudaMemcpyHostToDevice, ctx->stream));
    if (gc_class) CK(cudaMemcpyAsync(ctx->d_gc_class.p, gc_class, sizeof(int) * (size_t)J.n_seq * J.n_models, cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemsetAsync(ctx->d_gc_class.p, 0xff, sizeof(int) * (size_t)J.n_seq * J.n_models, ctx->stream));
    J.gc_log = (const double *)ctx->d_gc_log.p; J.gc_class = (const int *)ctx->d_gc_class.p;
    ctx->screens_dirty = true;
    return RMD_OK;
}

// Tabulate the prefix screens of every sequence with enough windows to repay it (the tables depend on the
// sequence's GC table, the cutoff and the models).  Issued on the scan stream in front of the first scan.
// Tabulating costs about as much as scanning a hundred windows and a block of tables is 66 KB for the shipped
// models, so short sequences go without (rmd_set_screen_policy), and the blocks of one job stay within a budget.
#define SCREEN_BUDGET_BYTES (256LL << 20)
static int build_screens(rmd_ctx *ctx) {
    DevJob &J = ctx->job;
    ctx->screens_dirty = false;
    J.screen_bits = nullptr; J.screen_off = nullptr; J.n_screens = ctx->n_screens;
    memcpy(J.screen0, ctx->screen0, sizeof J.screen0);
    if (ctx->n_screens == 0 || ctx->screen_words == 0) return RMD_OK;
    std::vector<long long> off((size_t)J.n_seq, -1);
    std::vector<int> seqs;
    const long long max_blocks = std::max<long long>(1, SCREEN_BUDGET_BYTES / (4 * ctx->screen_words));
    for (int s = 0; s < J.n_seq && (long long)seqs.size() < max_blocks; s++)
        if (ctx->screen_min_windows >= 0 && ctx->h_win_base[s + 1] - ctx->h_win_base[s] >= ctx->screen_min_windows) {
            off[s] = (long long)seqs.size() * ctx->screen_words;
            seqs.push_back(s);
        }
    if (seqs.empty()) return RMD_OK;
    RESERVE(ctx->d_screen_off, sizeof(long long) * (size_t)J.n_seq);
    RESERVE(ctx->d_screen_seqs, sizeof(int) * seqs.size());
    RESERVE(ctx->d_screen_bits, sizeof(uint32_t) * (size_t)ctx->screen_words * seqs.size());
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->d_screen_off.p, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_screen_seqs.p, seqs.data(), sizeof(int) * seqs.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                          // the host vectors go out of scope
    J.screen_bits = (const uint32_t *)ctx->d_screen_bits.p;
    J.screen_off = (const long long *)ctx->d_screen_off.p;
    const unsigned gx = (unsigned)(((1LL << (2 * SCREEN_MAX_LETTERS)) + 255) / 256);
    for (size_t z0 = 0; z0 < seqs.size(); z0 += 65535) {
        const unsigned gz = (unsigned)std::min<size_t>(65535, seqs.size() - z0);
        screen_build_kernel<<<dim3(gx, (unsigned)ctx->n_screens, gz), 256, 0, st>>>(J, (const int *)ctx->d_screen_seqs.p + z0, (uint32_t *)ctx->d_screen_bits.p);
    }
    CK(cudaGetLastError());
    return RMD_OK;
}

// Copy bpp entries [e0, e1) of a job to the device on `copy_st`; with the compact transport the lists the scan
// reads are rebuilt by expand_bpp_kernel, issued on the compute stream after the copy.
static int stage_entries(rmd_ctx *ctx, const rmd_scan_input *in, long long e0, long long e1, cudaStream_t copy_st, bool expand = true) {
    if (e1 <= e0) return RMD_OK;
    const size_t n = (size_t)(e1 - e0);
    if (in->bpp_q) {
        CK(cudaMemcpyAsync((uint16_t *)ctx->d_bpp_ij.p + e0, in->bpp_ij + e0, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, copy_st));
        CK(cudaMemcpyAsync((uint32_t *)ctx->d_bpp_q.p + e0, in->bpp_q + e0, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, copy_st));
        if (!expand) return RMD_OK;                         // scan_fast_kernel<..., FUSED> rebuilds the values window by window
        if (copy_st != ctx->stream) {
            CK(cudaEventRecord(ctx->ev[11], copy_st));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[11], 0));
        }
        expand_bpp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint16_t *)ctx->d_bpp_ij.p + e0, (const uint32_t *)ctx->d_bpp_q.p + e0, (long long)n,
                                                                                 (uint16_t *)ctx->d_bpp_i.p + e0, (uint16_t *)ctx->d_bpp_j.p + e0, (double *)ctx->d_bpp_p.p + e0);
    } else {
        CK(cudaMemcpyAsync((uint16_t *)ctx->d_bpp_i.p + e0, in->bpp_i + e0, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, copy_st));
        CK(cudaMemcpyAsync((uint16_t *)ctx->d_bpp_j.p + e0, in->bpp_j + e0, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, copy_st));
        CK(cudaMemcpyAsync((double *)ctx->d_bpp_p.p + e0, in->bpp_p + e0, sizeof(double) * n, cudaMemcpyHostToDevice, copy_st));
    }
    return RMD_OK;
}

// Stage a job.  with_entries == false leaves the bpp entry arrays (i, j, p) to the caller, which
// streams them in window chunks while earlier chunks are being scanned (rmd_scan).
static int upload_job(rmd_ctx *ctx, const rmd_scan_input *in, bool with_entries) {
    if (!ctx || !in) return fail(ctx, RMD_E_INVALID, "rmd_upload: NULL argument");
    if (ctx->n_models == 0) return fail(ctx, RMD_E_STATE, "rmd_upload: call rmd_set_models first");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    if (!in->gc_log) return fail(ctx, RMD_E_INVALID, "rmd_upload: gc_log is required (use rmd_gc_counts + rmd_set_gc to derive it on the device)");
    CK(cudaSetDevice(ctx->device));
    ctx->have_job = false;
    int r;
    if ((r = stage_layout(ctx, in->n_seq, in->seq_off, in->win_len, in->win_step, in->n_win))) return r;
    if ((r = stage_gc(ctx, in->gc_log, in->gc_class))) return r;
    DevJob &J = ctx->job;
    const long long n = ctx->n_letters;
    const long long nb = in->bpp_off ? in->bpp_off[in->n_win] : 0;
    ctx->n_bpp = nb;
    RESERVE(ctx->d_codes8, (size_t)n + 16);
    RESERVE(ctx->d_codes2, sizeof(uint32_t) * ((size_t)n / 16 + 2));
    RESERVE(ctx->d_flag, 64);
    RESERVE(ctx->d_bpp_off, sizeof(long long) * (in->n_win + 1));
    RESERVE(ctx->d_bpp_i, sizeof(uint16_t) * (size_t)std::max<long long>(nb, 1));
    RESERVE(ctx->d_bpp_j, sizeof(uint16_t) * (size_t)std::max<long long>(nb, 1));
    RESERVE(ctx->d_bpp_p, sizeof(double) * (size_t)std::max<long long>(nb, 1));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[4], st));
    if (n) CK(cudaMemcpyAsync(ctx->d_codes8.p, in->codes, (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_flag.p, 0, 64, st));
    if (n) {
        const long long words = (n + 15) / 16;
        pack_codes_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>((const uint8_t *)ctx->d_codes8.p, n, (uint32_t *)ctx->d_codes2.p, (int *)ctx->d_flag.p);
    }
    if (in->bpp_off) CK(cudaMemcpyAsync(ctx->d_bpp_off.p, in->bpp_off, sizeof(long long) * (in->n_win + 1), cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(ctx->d_bpp_off.p, 0, sizeof(long long) * (in->n_win + 1), st));
    const bool compact = in->bpp_q != nullptr;
    if (compact && !in->bpp_ij) return fail(ctx, RMD_E_INVALID, "rmd_upload: bpp_q needs bpp_ij");
    if (compact && ctx->max_wlen > 256) return fail(ctx, RMD_E_LIMIT, "compact bpp transport holds windows up to 256 letters (this job: %d)", ctx->max_wlen);
    if (nb && !compact && (!in->bpp_i || !in->bpp_j || !in->bpp_p)) return fail(ctx, RMD_E_INVALID, "rmd_upload: bpp_i / bpp_j / bpp_p are NULL");
    if (compact) {
        RESERVE(ctx->d_bpp_ij, sizeof(uint16_t) * (size_t)std::max<long long>(nb, 1));
        RESERVE(ctx->d_bpp_q, sizeof(uint32_t) * (size_t)std::max<long long>(nb, 1));
    }
    if (nb && with_entries) {
        int r2 = stage_entries(ctx, in, 0, nb, st);
        if (r2) return r2;
    }
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    ctx->any_other = flag != 0;
    J.codes2 = (const uint32_t *)ctx->d_codes2.p;
    J.codes8 = ctx->any_other ? (const uint8_t *)ctx->d_codes8.p : nullptr;
    J.bpp_off = (const long long *)ctx->d_bpp_off.p;
    J.bpp_i = (const uint16_t *)ctx->d_bpp_i.p; J.bpp_j = (const uint16_t *)ctx->d_bpp_j.p; J.bpp_p = (const double *)ctx->d_bpp_p.p;
    J.bpp_ij = nullptr; J.bpp_q = nullptr;
    J.min_bpp = in->min_bpp_mean; J.cutoff = in->cutoff; J.no_score = !(0.0 >= 0.0 + in->cutoff);
    job_thresholds(J);
    ctx->have_job = true;
    return RMD_OK;
}

extern "C" int rmd_upload(rmd_ctx *ctx, const rmd_scan_input *in) { return upload_job(ctx, in, true); }

extern "C" int rmd_gc_counts(rmd_ctx *ctx, int64_t *counts) {
    if (!ctx || !counts) return fail(ctx, RMD_E_INVALID, "rmd_gc_counts: NULL argument");
    if (!ctx->have_job) return fail(ctx, RMD_E_STATE, "rmd_gc_counts: no resident job");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    const DevJob &J = ctx->job;
    RESERVE(ctx->d_counts, sizeof(unsigned long long) * RMD_N_CODES * J.n_seq);
    CK(cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(unsigned long long) * RMD_N_CODES * J.n_seq, ctx->stream));
    long long longest = 0;
    for (int s = 0; s < J.n_seq; s++) longest = std::max(longest, ctx->h_seq_off[s + 1] - ctx->h_seq_off[s]);
    unsigned gx = (unsigned)std::min<long long>(std::max<long long>((longest + 256 * 16 - 1) / (256 * 16), 1), 148 * 8);
    for (int s0 = 0; s0 < J.n_seq; s0 += 65535) {
        DevJob part = J;
        part.seq_off = J.seq_off + s0;
        unsigned gy = (unsigned)std::min(65535, J.n_seq - s0);
        count_letters_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(part, (unsigned long long *)ctx->d_counts.p + (size_t)s0 * RMD_N_CODES);
    }
    CK(cudaMemcpyAsync(counts, ctx->d_counts.p, sizeof(unsigned long long) * RMD_N_CODES * J.n_seq, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return RMD_OK;
}

extern "C" int rmd_set_gc(rmd_ctx *ctx, const double *gc_log, const int32_t *gc_class) {
    if (!ctx || !gc_log) return fail(ctx, RMD_E_INVALID, "rmd_set_gc: NULL argument");
    if (!ctx->have_job) return fail(ctx, RMD_E_STATE, "rmd_set_gc: no resident job");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    int r = stage_gc(ctx, gc_log, gc_class);
    if (r) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return RMD_OK;
}

extern "C" int rmd_set_min_score(rmd_ctx *ctx, int enable, double min_score) {
    if (!ctx) return fail(ctx, RMD_E_INVALID, "rmd_set_min_score: NULL context");
    if (enable && !(min_score == min_score)) return fail(ctx, RMD_E_INVALID, "rmd_set_min_score: min_score is not a number");
    ctx->filter_score = enable != 0;
    ctx->min_score = min_score;
    return RMD_OK;
}

extern "C" int rmd_set_hit_format(rmd_ctx *ctx, int compact) {
    if (!ctx) return fail(ctx, RMD_E_INVALID, "rmd_set_hit_format: NULL context");
    ctx->compact_hits = compact != 0;
    return RMD_OK;
}

extern "C" int rmd_set_screen_policy(rmd_ctx *ctx, int64_t min_windows) {
    if (!ctx) return fail(ctx, RMD_E_INVALID, "rmd_set_screen_policy: NULL context");
    ctx->screen_min_windows = min_windows;
    ctx->screens_dirty = true;
    return RMD_OK;
}

extern "C" int rmd_set_return_dropped(rmd_ctx *ctx, int enable) {
    if (!ctx) return fail(ctx, RMD_E_INVALID, "rmd_set_return_dropped: NULL context");
    ctx->return_dropped = enable != 0;
    return RMD_OK;
}

// exclusive scan helper: out has n + 1 slots
static int device_scan(rmd_ctx *ctx, const uint32_t *in, long long n, uint32_t *out, int *launches) {
    if (n == 0) { CK(cudaMemsetAsync(out, 0, sizeof(uint32_t), ctx->stream)); return RMD_OK; }
    const int nblocks = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    RESERVE(ctx->d_scan_sums, sizeof(uint32_t) * (nblocks + 2));
    uint32_t *sums = (uint32_t *)ctx->d_scan_sums.p;
    scan_block_sums_kernel<<<nblocks, SCAN_BT, 0, ctx->stream>>>(in, n, sums);
    scan_sums_kernel<<<1, SCAN_BT, 0, ctx->stream>>>(sums, nblocks, sums + nblocks);
    scan_apply_kernel<<<nblocks, SCAN_BT, 0, ctx->stream>>>(in, n, sums, out);
    *launches += 3;
    return RMD_OK;
}

static int ensure_host(rmd_ctx *ctx, size_t hits, size_t groups) {
    if (hits > ctx->h_hits_cap) {
        if (ctx->h_hits) cudaFreeHost(ctx->h_hits);
        ctx->h_hits = nullptr; ctx->h_hits_cap = 0;
        size_t want = hits + hits / 4 + 1024;
        CK(cudaMallocHost((void **)&ctx->h_hits, want * sizeof(rmd_hit)));
        ctx->h_hits_cap = want;
    }
    if (groups > ctx->h_gate_cap) {
        if (ctx->h_gate) cudaFreeHost(ctx->h_gate);
        if (ctx->h_group) cudaFreeHost(ctx->h_group);
        ctx->h_gate = ctx->h_group = nullptr; ctx->h_gate_cap = 0;
        size_t want = groups + groups / 4 + 1024;
        CK(cudaMallocHost((void **)&ctx->h_gate, want * sizeof(uint32_t)));
        CK(cudaMallocHost((void **)&ctx->h_group, want * sizeof(uint32_t)));
        ctx->h_gate_cap = want;
    }
    return RMD_OK;
}

// Scan windows [w_begin, w_end) of the resident job.  When `stream_in` is given its bpp entry arrays
// are not on the device yet: they are copied in window chunks on the copy stream while the compute
// stream scans the chunks already there.
#define STREAM_CHUNKS 16
#ifndef STREAM_PIECES
#define STREAM_PIECES 80                  // copies of a streamed scan that runs as ONE launch (compact transport)
#endif
#ifndef STREAM_GROWTH_FUSED
#define STREAM_GROWTH_FUSED 45             // per cent by which a chunk's end grows over the previous one's (compact transport)
#endif
static int scan_core(rmd_ctx *ctx, int64_t w_begin, int64_t w_end, rmd_scan_result *out, const rmd_scan_input *stream_in) {
    if (!ctx || !out) return fail(ctx, RMD_E_INVALID, "rmd_scan_resident: NULL argument");
    if (!ctx->have_job) return fail(ctx, RMD_E_STATE, "rmd_scan_resident: no resident job");
    const DevJob &J = ctx->job;
    if (w_begin < 0 || w_end > J.n_win || w_begin > w_end) return fail(ctx, RMD_E_INVALID, "window range [%lld, %lld) outside [0, %lld)", (long long)w_begin, (long long)w_end, J.n_win);
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    { int rb = bind_tables(ctx); if (rb) return rb; }
    memset(out, 0, sizeof *out);
    const long long nwin = w_end - w_begin;
    const long long ngroups = nwin * J.n_models;
    if (ngroups >= (1LL << 31)) return fail(ctx, RMD_E_LIMIT, "window range too large for one call (%lld windows x %d models): stage the job with rmd_upload and scan it in ranges with rmd_scan_resident", nwin, J.n_models);
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[8], st));
    if (ctx->screens_dirty) { int rs = build_screens(ctx); if (rs) return rs; }
    RESERVE(ctx->d_counters, 128);
    RESERVE(ctx->d_next, sizeof(unsigned int) * (2 * STREAM_CHUNKS + 4));    // per chunk launch, OTHER launch, progress mark; the same for the big-window launches
    RESERVE(ctx->d_gate_pass, sizeof(uint32_t) * (size_t)std::max<long long>(ngroups, 1));
    RESERVE(ctx->d_group_hits, sizeof(uint32_t) * (size_t)std::max<long long>(ngroups, 1));
    RESERVE(ctx->d_group_off, sizeof(uint32_t) * (size_t)(ngroups + 1));
    if (ctx->raw_cap == 0) ctx->raw_cap = std::max<long long>(1 << 16, 24 * nwin);
    // the shared-memory kernel takes windows longer than every chain and up to 160 letters
    const bool use_fast = ctx->max_chain <= 32 && ctx->motif_ok;
    const int min_fast_wlen = use_fast ? ctx->max_chain + 1 : 0x7fffffff;
    // windows with a letter other than ACGU (N runs, IUPAC codes): a second launch of the shared-memory kernel takes them
    // from a work list; the rare kernel shapes (exact gate, local-memory stack, tables in global memory) leave them to the generic kernel
    const bool other_fast = use_fast && ctx->any_other && ctx->tables_in_smem && !ctx->local_stack && !(ctx->any_brute || !(J.min_bpp > 0.0));
    // windows of 161..256 letters (a larger --win-len, whole-sequence mode on short sequences): a second launch of the kernel staged for them
    const bool big_fast = use_fast && ctx->max_wlen > RMD_FAST_WINDOW && ctx->teams_big > 0 && ctx->tables_in_smem && !ctx->local_stack &&
                          !(ctx->any_brute || !(J.min_bpp > 0.0));
    const int max_fast_wlen = big_fast ? RMD_BIG_WINDOW : RMD_FAST_WINDOW;
    const bool need_generic = !use_fast || ctx->max_wlen > max_fast_wlen || ctx->min_wlen < min_fast_wlen || (ctx->any_other && !other_fast);
    if (other_fast) RESERVE(ctx->d_other, sizeof(uint32_t) * (size_t)(nwin + 2));
    unsigned long long counters[16];
    int launches = 0;
    bool grown = false;                              // the record buffer has been enlarged once for this call
    for (int attempt = 0;; attempt++) {
        RESERVE(ctx->d_raw, sizeof(RawHit) * (size_t)ctx->raw_cap);
        launches = 0;
        CK(cudaMemsetAsync(ctx->d_counters.p, 0, 128, st));
        CK(cudaMemsetAsync(ctx->d_next.p, 0, sizeof(unsigned int) * (2 * STREAM_CHUNKS + 4), st));
        CK(cudaMemsetAsync(ctx->d_gate_pass.p, 0, sizeof(uint32_t) * (size_t)std::max<long long>(ngroups, 1), st));
        CK(cudaMemsetAsync(ctx->d_group_hits.p, 0, sizeof(uint32_t) * (size_t)std::max<long long>(ngroups, 1), st));
        ScanOut O;
        O.raw = (RawHit *)ctx->d_raw.p; O.raw_count = (unsigned long long *)ctx->d_counters.p; O.stats = O.raw_count + 8; O.raw_cap = ctx->raw_cap;
        O.gate_pass = (uint32_t *)ctx->d_gate_pass.p; O.group_hits = (uint32_t *)ctx->d_group_hits.p;
        O.other_list = (const uint32_t *)ctx->d_other.p; O.other_count = O.other_list ? O.other_list + nwin : nullptr;
        O.ready = nullptr;
        CK(cudaEventRecord(ctx->ev[0], st));
        if (other_fast && nwin > 0) {
            CK(cudaMemsetAsync((uint32_t *)ctx->d_other.p + nwin, 0, sizeof(uint32_t), st));
            other_windows_kernel<<<(unsigned)std::min<long long>((nwin + 7) / 8, (long long)ctx->sm_count * 8), 256, 0, st>>>(J, w_begin, w_end, min_fast_wlen, max_fast_wlen, (uint32_t *)ctx->d_other.p,
                                                                                                                            (uint32_t *)ctx->d_other.p + nwin);
            launches++;
        }
        // Chunks of windows whose lists are copied while earlier chunks are scanned.  Every chunk is one launch of
        // the persistent kernel and pays its tail (0.1 ms), and the first must be small so that the scan starts
        // early.  A chunk can start once its copy is complete: with the copy r times faster per window than the
        // scan, the scan never waits if every chunk END grows by at most the factor r over the previous one
        // (compact transport on a B200: copy 10 ms, scan 13 ms per 133 k windows, r = 1.3; measured best: ends growing
        // by 1.6, or by 1.45 when the launches overlap — a launch more costs more than a short wait).  Doubling chunk SIZES, as an earlier version did,
        // left the last chunk (half of the data) waiting 1.7 ms for its copy.
        const bool streaming = stream_in && attempt == 0 && ctx->n_bpp > 0;
        long long bounds[STREAM_CHUNKS + 1];
        int n_chunks = 1;
        bounds[0] = w_begin; bounds[1] = w_end;
        if (streaming && nwin >= 4096) {
            if (stream_in->bpp_q) {
                n_chunks = 0;
                // without expansion launches, and with the chunks' tails overlapping on two streams, a launch costs less
                const bool will_fuse = use_fast && !need_generic && !big_fast && !ctx->any_other && ctx->tables_in_smem && !ctx->local_stack &&
                                       !(ctx->any_brute || !(J.min_bpp > 0.0)) && !tune_env("RMD_NO_FUSED_EXPAND");
                const int gpct = will_fuse ? STREAM_GROWTH_FUSED : 60;
                long long end = std::max<long long>(nwin / 128, 2048);
                while (end < nwin && n_chunks < STREAM_CHUNKS - 1) { bounds[++n_chunks] = w_begin + end; end = end + end * gpct / 100 + 1024; }
                bounds[++n_chunks] = w_end;
            } else {
                n_chunks = (int)std::min<long long>(STREAM_CHUNKS, nwin / 2048);
                for (int k = 0; k <= n_chunks; k++) bounds[k] = w_begin + nwin * k / n_chunks;
            }
        }
        // Compact transport and every window the shared-memory kernel's: no expansion launches between the scan launches,
        // the teams rebuild the values of their own windows (and keep doing so while the job stays resident).  The
        // chunks then alternate between two compute streams: a launch does not wait for the previous one to end, its
        // CTAs start on the SMs whose persistent CTA has run out of windows.
        if (streaming) {
            const bool fused = stream_in->bpp_q && use_fast && !need_generic && !big_fast && !ctx->any_other && ctx->tables_in_smem && !ctx->local_stack &&
                               !(ctx->any_brute || !(J.min_bpp > 0.0)) && !tune_env("RMD_NO_FUSED_EXPAND");     // = will_fuse above
            ctx->job.bpp_ij = fused ? (const uint16_t *)ctx->d_bpp_ij.p : nullptr;
            ctx->job.bpp_q = fused ? (const uint32_t *)ctx->d_bpp_q.p : nullptr;
        }
        const bool fused_expand = ctx->job.bpp_q != nullptr;
        const bool two_streams = false;                            // (chunk launches of the compact transport alternated between two streams before the one-launch scan)
        if (two_streams) {                                          // the second compute stream starts after the resets above
            CK(cudaEventRecord(ctx->ev[12], st));
            CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev[12], 0));
        }
        const cudaStream_t first_stream = st;
        // Compact transport, every window the shared-memory kernel's: ONE launch scans the whole range while the copy
        // stream is still delivering it.  After every piece of the lists the copy stream writes a progress mark (the
        // number of windows that are complete); a team that draws a window beyond the mark waits for it.  No launch
        // tails between chunks, no chunk schedule to tune: the scan ends when the slower of copy and kernel ends.
        const bool one_launch = streaming && fused_expand && nwin > 0 && n_chunks > 1;
        if (one_launch) {
            if (!ctx->h_ready) CK(cudaMallocHost((void **)&ctx->h_ready, sizeof(unsigned int) * (STREAM_PIECES + 2)));
            unsigned int *d_ready = (unsigned int *)ctx->d_next.p + STREAM_CHUNKS + 1;
            O.ready = d_ready;
            CK(cudaEventRecord(ctx->ev[12], st));                   // the mark was zeroed on `st`: marks come after that
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[12], 0));
            const unsigned grid = (unsigned)std::min<long long>((long long)ctx->sm_count * ctx->ctas_per_sm, (nwin + ctx->teams - 1) / ctx->teams);
            scan_fast_kernel<false, true, false, true><<<grid, (unsigned)ctx->teams * SCAN_THREADS, ctx->scan_smem, st>>>(
                J, w_begin, w_end, w_begin, O, min_fast_wlen, (unsigned int *)ctx->d_next.p);
            launches++;
            // a small first piece (the kernel starts on it), equal pieces after it
            const long long first = std::min<long long>(nwin, std::max<long long>(nwin / 128, 1024));
            const long long rest = (nwin - first + STREAM_PIECES - 2) / (STREAM_PIECES - 1);
            int rc = RMD_OK, k = 0;
            for (long long c0 = 0; c0 < nwin && rc == RMD_OK; k++) {
                const long long c1 = k == 0 ? first : std::min(nwin, c0 + std::max<long long>(rest, 1));
                rc = stage_entries(ctx, stream_in, stream_in->bpp_off[w_begin + c0], stream_in->bpp_off[w_begin + c1], ctx->copy_stream, false);
                ctx->h_ready[k] = (unsigned int)c1;
                if (rc == RMD_OK && cudaMemcpyAsync(d_ready, ctx->h_ready + k, sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess)
                    rc = fail(ctx, RMD_E_CUDA, "progress mark of a streamed scan could not be enqueued: %s", cudaGetErrorString(cudaGetLastError()));
                c0 = c1;
            }
            if (rc != RMD_OK) {                                     // never leave the kernel waiting for lists that will not come
                ctx->h_ready[STREAM_PIECES + 1] = 0xffffffffu;
                cudaMemcpyAsync(d_ready, ctx->h_ready + STREAM_PIECES + 1, sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->copy_stream);
                cudaStreamSynchronize(ctx->copy_stream);
                cudaStreamSynchronize(st);
                return rc;
            }
        }
        for (int ch = 0; ch < n_chunks && nwin > 0 && !one_launch; ch++) {
            const long long c0 = bounds[ch], c1 = bounds[ch + 1];
            const cudaStream_t st = two_streams && (ch & 1) ? ctx->stream2 : first_stream;
            if (streaming) {                                        // entries of windows [c0, c1): copy stream, then let the scan wait for them
                const long long e0 = stream_in->bpp_off[c0], e1 = stream_in->bpp_off[c1];
                int r3 = stage_entries(ctx, stream_in, e0, e1, ctx->copy_stream, !fused_expand);
                if (r3) return r3;
                if (stream_in->bpp_q && e1 > e0 && !fused_expand) launches++;
                CK(cudaEventRecord(ctx->ev[10], ctx->copy_stream));
                CK(cudaStreamWaitEvent(st, ctx->ev[10], 0));
            }
            if (c1 <= c0) continue;
            if (use_fast) {
                unsigned int *next = (unsigned int *)ctx->d_next.p + ch;      // this launch's window counter
                const bool brute = ctx->any_brute || !(J.min_bpp > 0.0);
                const unsigned grid = (unsigned)std::min<long long>((long long)ctx->sm_count * ctx->ctas_per_sm, (c1 - c0 + ctx->teams - 1) / ctx->teams);
                const unsigned block = (unsigned)ctx->teams * SCAN_THREADS;
                if (ctx->tables_in_smem) {
                    // the exact-gate and global-table variants are the rare ones: they keep the local-memory stack
                    if (brute) scan_fast_kernel<true, true, true><<<grid, block, ctx->scan_smem, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, next);
                    else if (ctx->local_stack) scan_fast_kernel<false, true, true><<<grid, block, ctx->scan_smem, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, next);
                    else if (fused_expand) scan_fast_kernel<false, true, false, true><<<grid, block, ctx->scan_smem, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, next);
                    else scan_fast_kernel<false, true, false><<<grid, block, ctx->scan_smem, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, next);
                } else {
                    if (brute) scan_fast_kernel<true, false, true><<<grid, block, ctx->scan_smem, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, next);
                    else scan_fast_kernel<false, false, true><<<grid, block, ctx->scan_smem, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, next);
                }
                launches++;
            }
            if (big_fast) {                                         // windows of 161..256 letters of this chunk
                const unsigned grid = (unsigned)std::min<long long>((long long)ctx->sm_count * ctx->ctas_per_sm, (c1 - c0 + ctx->teams_big - 1) / ctx->teams_big);
                scan_fast_kernel<false, true, false, false, false, RMD_BIG_WINDOW><<<grid, (unsigned)ctx->teams_big * SCAN_THREADS, ctx->scan_smem_big, st>>>(
                    J, c0, c1, w_begin, O, RMD_FAST_WINDOW + 1, (unsigned int *)ctx->d_next.p + STREAM_CHUNKS + 2 + ch);
                launches++;
            }
            if (need_generic) {
                scan_generic_kernel<<<(unsigned)((c1 - c0) * J.n_models), GEN_THREADS, 0, st>>>(J, c0, c1, w_begin, O, min_fast_wlen, max_fast_wlen, other_fast ? 1 : 0);
                launches++;
            }
        }
        if (two_streams) {                                          // join
            CK(cudaEventRecord(ctx->ev[13], ctx->stream2));
            CK(cudaStreamWaitEvent(st, ctx->ev[13], 0));
        }
        if (other_fast && nwin > 0) {                               // the windows the launches above skipped; every list is on the device by now
            scan_fast_kernel<false, true, false, false, true><<<(unsigned)ctx->sm_count * ctx->ctas_per_sm, (unsigned)ctx->teams * SCAN_THREADS, ctx->scan_smem, st>>>(
                J, w_begin, w_end, w_begin, O, min_fast_wlen, (unsigned int *)ctx->d_next.p + STREAM_CHUNKS);
            launches++;
            if (big_fast) {
                scan_fast_kernel<false, true, false, false, true, RMD_BIG_WINDOW><<<(unsigned)ctx->sm_count * ctx->ctas_per_sm, (unsigned)ctx->teams_big * SCAN_THREADS, ctx->scan_smem_big, st>>>(
                    J, w_begin, w_end, w_begin, O, RMD_FAST_WINDOW + 1, (unsigned int *)ctx->d_next.p + 2 * STREAM_CHUNKS + 3);
                launches++;
            }
        }
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaMemcpyAsync(counters, ctx->d_counters.p, 128, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (counters[7] & ERR_BAD_LIST) return fail(ctx, RMD_E_INVALID, "a window's bpp list is not a sorted, duplicate-free upper triangle (i < j < window length)");
        if (counters[7] & ERR_BAD_VALUE) return fail(ctx, RMD_E_INVALID, "a bpp value is negative or not a number");
        if (counters[7] & ERR_TIMEOUT) {
            // the copies did not advance while the launch was waiting for them: something serialises the GPU's work (ncu,
            // a debugger).  The pieces are all enqueued; wait for them and scan the resident lists with a plain launch.
            CK(cudaStreamSynchronize(ctx->copy_stream));
            if (one_launch && attempt == 0) { ctx->stream_fallbacks++; continue; }
            return fail(ctx, RMD_E_CUDA, "streamed scan: the lists of some windows did not arrive");
        }
        // Records are appended as they are found, their number is known only afterwards (at most one per gate pass).  The
        // buffer holds 24 per window — five times what the shipped models leave on random sequence — and keeps the size it
        // last needed; a scan that outgrows it is repeated once with room for exactly the number just counted.
        if ((long long)counters[0] <= ctx->raw_cap) break;
        if (grown) return fail(ctx, RMD_E_NOMEM, "hit buffer overflow persisted (%llu records)", counters[0]);
        grown = true;
        ctx->raw_cap = (long long)counters[0] + (long long)counters[0] / 8 + 1024;   // rerun with room for every record
    }
    const long long n_raw = (long long)counters[0];
    out->n_raw = n_raw;
    out->tries[0] = (int64_t)counters[1];
    out->tries[1] = (int64_t)counters[2];
#ifdef RMD_STATS
    fprintf(stderr, "[rmd stats] tree visits %llu in %llu warp iterations; warp cycles: staging %llu, generate+gate+score %llu, "
            "window-end engine %llu, barrier+tail %llu; candidate funnel: %llu (entry, node) pairs -> %llu with all ancestors -> %llu new placements (%llu sure)\n",
            counters[4], counters[5], counters[8], counters[9], counters[10], counters[11], counters[12], counters[13], counters[14], counters[15]);
#endif
    // ---- ordering, scoring, duplicate removal, compaction ---------------------------------------
    RESERVE(ctx->d_final, sizeof(rmd_hit) * (size_t)std::max<long long>(n_raw, 1));
    RESERVE(ctx->d_compact, sizeof(rmd_hit) * (size_t)std::max<long long>(n_raw, 1));
    RESERVE(ctx->d_keep, sizeof(uint32_t) * (size_t)std::max<long long>(n_raw, 1));
    RESERVE(ctx->d_keep_off, sizeof(uint32_t) * (size_t)(n_raw + 1));
    CK(cudaEventRecord(ctx->ev[2], st));
    int r;
    if ((r = device_scan(ctx, (const uint32_t *)ctx->d_group_hits.p, ngroups, (uint32_t *)ctx->d_group_off.p, &launches))) return r;
    uint32_t n_keep = 0;
    if (n_raw > 0) {
        const unsigned blocks = (unsigned)((n_raw + 255) / 256);
        // d_keep doubles as the per-group fill counters, d_compact as the bucketed record list
        RESERVE(ctx->d_keep, sizeof(uint32_t) * (size_t)std::max<long long>(std::max(n_raw, ngroups), 1));
        CK(cudaMemsetAsync(ctx->d_keep.p, 0, sizeof(uint32_t) * (size_t)ngroups, st));
        place_hits_kernel<<<blocks, 256, 0, st>>>((const RawHit *)ctx->d_raw.p, n_raw, J.n_models, (const uint32_t *)ctx->d_group_off.p,
                                                  (uint32_t *)ctx->d_keep.p, (RawHit *)ctx->d_compact.p);
        finalize_hits_kernel<<<blocks, 256, 0, st>>>(J, w_begin, (const RawHit *)ctx->d_compact.p, n_raw, (const uint32_t *)ctx->d_group_off.p,
                                                     std::log(2.0), (rmd_hit *)ctx->d_final.p);
        launches++;
        dedupe_kernel<<<blocks, 256, 0, st>>>(J, w_begin, w_end, (rmd_hit *)ctx->d_final.p, n_raw, (const uint32_t *)ctx->d_group_off.p,
                                              (uint32_t *)ctx->d_keep.p, ctx->filter_score ? 1 : 0, ctx->min_score);
        launches += 2;
        if ((r = device_scan(ctx, (const uint32_t *)ctx->d_keep.p, n_raw, (uint32_t *)ctx->d_keep_off.p, &launches))) return r;
        compact_hits_kernel<<<blocks, 256, 0, st>>>((const rmd_hit *)ctx->d_final.p, n_raw, (const uint32_t *)ctx->d_keep.p,
                                                    (const uint32_t *)ctx->d_keep_off.p, (rmd_hit *)ctx->d_compact.p);
        launches++;
        CK(cudaMemcpyAsync(&n_keep, (uint32_t *)ctx->d_keep_off.p + n_raw, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const long long n_out = ctx->return_dropped ? n_raw : (long long)n_keep;
    ctx->last_hits = (const rmd_hit *)ctx->d_compact.p; ctx->last_n = (long long)n_keep;      // rmd_table_append_scan
    if ((r = ensure_host(ctx, (size_t)n_out, (size_t)ngroups))) return r;
    CK(cudaEventRecord(ctx->ev[6], st));
    const void *d_list = ctx->return_dropped ? ctx->d_final.p : ctx->d_compact.p;
    size_t hit_bytes = sizeof(rmd_hit);
    if (ctx->compact_hits && n_out) {
        RESERVE(ctx->d_pack, sizeof(rmd_hit32) * (size_t)n_out);
        pack_hits32_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>((const rmd_hit *)d_list, n_out, (rmd_hit32 *)ctx->d_pack.p);
        launches++;
        d_list = ctx->d_pack.p;
    }
    if (ctx->compact_hits) hit_bytes = sizeof(rmd_hit32);
    if (n_out) CK(cudaMemcpyAsync(ctx->h_hits, d_list, hit_bytes * n_out, cudaMemcpyDeviceToHost, st));
    if (ngroups) {
        CK(cudaMemcpyAsync(ctx->h_gate, ctx->d_gate_pass.p, sizeof(uint32_t) * ngroups, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(ctx->h_group, ctx->d_group_hits.p, sizeof(uint32_t) * ngroups, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(ctx->ev[7], st));
    CK(cudaEventRecord(ctx->ev[9], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&out->ms_total, ctx->ev[8], ctx->ev[9]));
    out->n_hits = n_out;
    out->hits = ctx->h_hits;
    out->hit_bytes = (int32_t)hit_bytes;
    out->gate_pass = ctx->h_gate;
    out->raw_count = ctx->h_group;
    CK(cudaEventElapsedTime(&out->ms_scan, ctx->ev[0], ctx->ev[1]));
    CK(cudaEventElapsedTime(&out->ms_post, ctx->ev[2], ctx->ev[3]));
    CK(cudaEventElapsedTime(&out->ms_d2h, ctx->ev[6], ctx->ev[7]));
    out->n_launches = launches;
    return RMD_OK;
}

extern "C" int rmd_scan_resident(rmd_ctx *ctx, int64_t w_begin, int64_t w_end, rmd_scan_result *out) {
    return scan_core(ctx, w_begin, w_end, out, nullptr);
}

extern "C" int rmd_scan(rmd_ctx *ctx, const rmd_scan_input *in, rmd_scan_result *out) {
    if (!ctx || !in || !out) return fail(ctx, RMD_E_INVALID, "rmd_scan: NULL argument");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    int r = upload_job(ctx, in, false);
    if (r) return r;
    // the copy stream must not start before the staging above has been issued ... it has been synchronised
    r = scan_core(ctx, 0, ctx->job.n_win, out, in);
    if (r) {                                       // the lists may be on the device only in part: the job is not resident
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream2);
        ctx->have_job = false;
        return r;
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    out->ms_h2d = ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[9]));
    out->ms_total = ms;
    out->n_launches += 1;
    return RMD_OK;
}

extern "C" int rmd_eval(rmd_ctx *ctx, int model, const uint8_t *codes, const int64_t *seq_off, const double *gc_log,
                        const int32_t *p0, const int32_t *p1, int64_t n, double cutoff,
                        int32_t *leaf, double *prob_model, double *prob_rand) {
    if (!ctx || model < 0 || model >= ctx->n_models) return fail(ctx, RMD_E_INVALID, "rmd_eval: bad model index");
    if (n <= 0) return RMD_OK;
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    { int rb = bind_tables(ctx); if (rb) return rb; }
    cudaStream_t st = ctx->stream;
    ctx->cal_open = false;                         // the scratch buffers below are a calibration session's too
    const long long nl = seq_off[n];
    const CModel &M = ctx->hmodels[model];
    for (int64_t k = 0; k < n; k++) {            // the reference would raise IndexError past the end
        const long long len = seq_off[k + 1] - seq_off[k];
        if (p0[k] < 0 || p1[k] < 0 || p0[k] + M.chain_len[0] > len || p1[k] + M.chain_len[1] > len)
            return fail(ctx, RMD_E_INVALID, "rmd_eval: placement %lld reaches outside its sequence", (long long)k);
    }
    RESERVE(ctx->d_e0, (size_t)nl + 16);
    RESERVE(ctx->d_e1, sizeof(long long) * (n + 1));
    RESERVE(ctx->d_e2, sizeof(double) * RMD_N_CODES * n);
    RESERVE(ctx->d_e3, sizeof(int) * n);
    RESERVE(ctx->d_e4, sizeof(int) * n);
    RESERVE(ctx->d_e5, sizeof(int) * n);
    RESERVE(ctx->d_e6, sizeof(double) * n);
    RESERVE(ctx->d_e7, sizeof(double) * n);
    CK(cudaMemcpyAsync(ctx->d_e0.p, codes, (size_t)nl, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_e1.p, seq_off, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_e2.p, gc_log, sizeof(double) * RMD_N_CODES * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_e3.p, p0, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_e4.p, p1, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    eval_batch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(model, (const uint8_t *)ctx->d_e0.p, (const long long *)ctx->d_e1.p,
                                                                   (const double *)ctx->d_e2.p, (const int *)ctx->d_e3.p, (const int *)ctx->d_e4.p,
                                                                   n, cutoff, (int *)ctx->d_e5.p, (double *)ctx->d_e6.p, (double *)ctx->d_e7.p);
    CK(cudaMemcpyAsync(leaf, ctx->d_e5.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(prob_model, ctx->d_e6.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(prob_rand, ctx->d_e7.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return RMD_OK;
}

// ---- calibration (rmcalibrate.py:235-298): a session keeps the weighted histograms on the device ---------------
// rmd_calib_open stages the class logs and rounding thresholds and zeroes the histogram; rmd_calib_add enqueues the
// scoring of a batch of samples (no synchronisation: batches queue up behind each other); rmd_calib_device_hist hands
// out the histogram's device address so that several GPUs can all-reduce it in place; rmd_calib_read downloads it.
static int calib_thresholds(std::vector<double> &thr, int n_bins) {
    thr.resize(n_bins);
    for (int k = 0; k < n_bins; k++) {
        // smallest double whose decimal rounding to one place (Python's round(x, 1)) reaches (k+1)/10: probe around the midpoint
        double mid = (2.0 * k + 1.0) / 20.0, lo = nextafter(mid, 0.0), best = nextafter(mid, 1e300);
        char a[32];
        for (double c : {lo, mid}) {
            snprintf(a, sizeof a, "%.1f", c);
            if (lround(strtod(a, nullptr) * 10.0) >= k + 1) { best = c; break; }
        }
        thr[k] = best;
    }
    return RMD_OK;
}

extern "C" int rmd_calib_open(rmd_ctx *ctx, int model, int32_t L, const double *class_log, int32_t n_class, double cutoff, int32_t n_bins) {
    if (!ctx || model < 0 || model >= ctx->n_models) return fail(ctx, RMD_E_INVALID, "rmd_calib_open: bad model index");
    const CModel &M = ctx->hmodels[model];
    if (L != M.chain_len[0] + M.chain_len[1]) return fail(ctx, RMD_E_INVALID, "rmd_calib_open: L is %d, the model has %d positions", L, M.chain_len[0] + M.chain_len[1]);
    if (L > 32) return fail(ctx, RMD_E_LIMIT, "rmd_calib_open: at most 32 positions");
    if (!class_log || n_class < 1 || n_class > CAL_MAX_CLASS || n_bins < 2) return fail(ctx, RMD_E_LIMIT, "rmd_calib_open: class count %d / bin count %d out of range", n_class, n_bins);
    // a weight is a product of powers of frequency ratios (calib_kernels.cuh): every frequency must be positive, as in the
    // reference, whose math.log raises on a zero (rmcalibrate.py:224, lib/gcx.py)
    for (int k = 0; k < 4 * n_class; k++)
        if (!std::isfinite(class_log[k])) return fail(ctx, RMD_E_INVALID, "rmd_calib_open: class %d has a log-frequency that is not finite", k / 4);
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n_cells = (size_t)(L + 1) * (L + 1) * (L + 1);
    ctx->cal_smem = L <= CAL_SMEM_L ? sizeof(uint32_t) * n_cells : 0;
    RESERVE(ctx->d_caltab, sizeof(double) * (size_t)n_class * 3 * (L + 1));
    RESERVE(ctx->d_calcells, sizeof(uint32_t) * (n_cells + 1));
    RESERVE(ctx->d_calcls, sizeof(double) * 4 * n_class);
    RESERVE(ctx->d_calhist, sizeof(double) * (size_t)n_class * n_bins);
    RESERVE(ctx->d_calcnt, 64);
    if (ctx->cal_thr_bins != n_bins) {               // the rounding thresholds depend on the bucket count only: once per context
        std::vector<double> thr;
        calib_thresholds(thr, n_bins);
        ctx->cal_thr_bins = 0;
        RESERVE(ctx->d_calthr, sizeof(double) * n_bins);
        CK(cudaMemcpy(ctx->d_calthr.p, thr.data(), sizeof(double) * n_bins, cudaMemcpyHostToDevice));
        ctx->cal_thr_bins = n_bins;
    }
    // the class logs leave through a pinned copy of the context's: no wait for the caller's buffer
    if (!ctx->h_calcls) CK(cudaMallocHost((void **)&ctx->h_calcls, sizeof(double) * 4 * CAL_MAX_CLASS));
    else CK(cudaStreamSynchronize(st));              // a previous session's copy may still be in flight (it never is after a read)
    memcpy(ctx->h_calcls, class_log, sizeof(double) * 4 * n_class);
    CK(cudaMemcpyAsync(ctx->d_calcls.p, ctx->h_calcls, sizeof(double) * 4 * n_class, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_calhist.p, 0, sizeof(double) * (size_t)n_class * n_bins, st));
    CK(cudaMemsetAsync(ctx->d_calcnt.p, 0, 64, st));
    calib_table_kernel<<<(unsigned)((n_class * 3 * (L + 1) + 255) / 256), 256, 0, st>>>((const double *)ctx->d_calcls.p, n_class, L, (double *)ctx->d_caltab.p);
    CK(cudaGetLastError());
    CalibArgs &A = ctx->cal;
    A.model = model; A.L = L; A.n_class = n_class; A.n_bins = n_bins; A.n = 0; A.first = 0;
    A.samples = nullptr; A.seed = 0;
    A.class_log = (const double *)ctx->d_calcls.p; A.thr = (const double *)ctx->d_calthr.p;
    A.wtab = (const double *)ctx->d_caltab.p; A.cells = (uint32_t *)ctx->d_calcells.p;
    A.cutoff = cutoff; A.ln2 = std::log(2.0); A.K = L * std::log(4.0);
    A.hist = (double *)ctx->d_calhist.p;
    A.n_ok = (unsigned long long *)ctx->d_calcnt.p; A.overflow = (int *)((char *)ctx->d_calcnt.p + 8);
    ctx->cal_open = true; ctx->cal_timed = false;
    return RMD_OK;
}

extern "C" int rmd_calib_add(rmd_ctx *ctx, const uint8_t *samples, uint64_t seed, int64_t first, int64_t n) {
    if (!ctx || !ctx->cal_open) return fail(ctx, RMD_E_STATE, "rmd_calib_add: call rmd_calib_open first");
    if (n <= 0) return RMD_OK;
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    { int rb = bind_tables(ctx); if (rb) return rb; }
    cudaStream_t st = ctx->stream;
    CalibArgs A = ctx->cal;
    if (samples) {
        CK(cudaStreamSynchronize(st));             // the previous batch may still be reading the staging buffer
        RESERVE(ctx->d_e0, (size_t)n * A.L);
        CK(cudaMemcpyAsync(ctx->d_e0.p, samples, (size_t)n * A.L, cudaMemcpyHostToDevice, st));
        A.samples = (const uint8_t *)ctx->d_e0.p;
    }
    A.seed = seed;
    if (!ctx->cal_timed) { CK(cudaEventRecord(ctx->ev[0], st)); ctx->cal_timed = true; }
    // the scoring engine's node format holds the model set (chains up to 16 letters) and the root passes its own test
    // (lib/node.py:248: 0.0 >= 0.0 + cutoff): lane-level walks; else one sample per thread through eval_placement
    const bool engine = ctx->motif_ok && ctx->hmodels[A.model].n_tree > 0 && 0.0 >= 0.0 + A.cutoff && !tune_env("RMD_CAL_NO_ENGINE");
    A.xnodes = (const XNode *)ctx->d_xnodes.p; A.xcpt = (const double *)ctx->d_xcpt.p;
    // a batch goes through in pieces of CAL_PIECE samples: the list of solved samples is sized for a piece (every sample
    // of it may have a solution) and a count cell cannot overflow
    const long long piece = std::min<long long>(n, CAL_PIECE);
    RESERVE(ctx->d_calok, sizeof(CalibOk) * (size_t)piece);
    A.ok = (CalibOk *)ctx->d_calok.p;
    const uint8_t *const staged = A.samples;
    for (long long c0 = 0; c0 < n; c0 += piece) {
        const long long cn = std::min(piece, n - c0);
        A.n = cn; A.first = first + c0;
        if (staged) A.samples = staged + (size_t)c0 * A.L;
        // persistent CTAs: the count histogram is flushed once per CTA
        const unsigned grid = (unsigned)std::min<long long>((cn + CAL_THREADS - 1) / CAL_THREADS, (long long)ctx->sm_count * CAL_CTAS_PER_SM);
        CK(cudaMemsetAsync(A.cells, 0, sizeof(uint32_t) * ((size_t)(A.L + 1) * (A.L + 1) * (A.L + 1) + 1), st));
        if (engine) {
            if (ctx->cal_smem) calibrate_engine_kernel<true><<<grid, CAL_THREADS, ctx->cal_smem, st>>>(A);
            else calibrate_engine_kernel<false><<<grid, CAL_THREADS, 0, st>>>(A);
        } else {
            if (ctx->cal_smem) calibrate_kernel<true><<<grid, CAL_THREADS, ctx->cal_smem, st>>>(A);
            else calibrate_kernel<false><<<grid, CAL_THREADS, 0, st>>>(A);
        }
        calib_ok_kernel<<<(unsigned)std::min<long long>((cn * A.n_class + 255) / 256, (long long)ctx->sm_count * 8), 256, 0, st>>>(A);
        calib_cells_kernel<<<(unsigned)A.n_class, 256, 0, st>>>(A);
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaGetLastError());
    if (samples) CK(cudaStreamSynchronize(st));    // the caller's buffer may be reused on return
    return RMD_OK;
}

extern "C" int rmd_calib_device_hist(rmd_ctx *ctx, void **dptr, int64_t *n_doubles) {
    if (!ctx || !dptr || !ctx->cal_open) return fail(ctx, RMD_E_STATE, "rmd_calib_device_hist: no open calibration session");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));        // the batches enqueued so far are in the histogram
    CK(cudaGetLastError());
    *dptr = ctx->cal.hist;
    if (n_doubles) *n_doubles = (int64_t)ctx->cal.n_class * ctx->cal.n_bins;
    return RMD_OK;
}

extern "C" int rmd_calib_read(rmd_ctx *ctx, double *hist, int64_t *n_ok, int32_t *overflow, float *ms_device) {
    if (!ctx || !ctx->cal_open) return fail(ctx, RMD_E_STATE, "rmd_calib_read: no open calibration session");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    unsigned long long small[2] = {0, 0};
    if (hist) CK(cudaMemcpyAsync(hist, ctx->cal.hist, sizeof(double) * (size_t)ctx->cal.n_class * ctx->cal.n_bins, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(small, ctx->d_calcnt.p, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (n_ok) *n_ok = (int64_t)small[0];
    if (overflow) *overflow = (int32_t)(small[1] & 0xffffffffu);
    if (ms_device) { *ms_device = 0.0f; if (ctx->cal_timed) CK(cudaEventElapsedTime(ms_device, ctx->ev[0], ctx->ev[1])); }
    ctx->cal_timed = false;
    return RMD_OK;
}

extern "C" int rmd_calib_read_rows(rmd_ctx *ctx, int32_t first_class, int32_t n_rows, double *rows, void **dptr, int32_t *cols) {
    if (!ctx || !ctx->cal_open) return fail(ctx, RMD_E_STATE, "rmd_calib_read_rows: no open calibration session");
    if (first_class < 0 || n_rows < 1 || first_class + n_rows > ctx->cal.n_class) return fail(ctx, RMD_E_INVALID, "rmd_calib_read_rows: classes %d..%d of %d", first_class, first_class + n_rows, ctx->cal.n_class);
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const double *src = ctx->cal.hist + (size_t)first_class * ctx->cal.n_bins;
    int range[2] = {ctx->cal.n_bins, -1};
    int *d_range = (int *)((char *)ctx->d_calcnt.p + 16);          // behind n_ok and the overflow flag
    if (cols) {
        CK(cudaMemcpyAsync(d_range, range, sizeof range, cudaMemcpyHostToDevice, st));
        calib_cols_kernel<<<(unsigned)ctx->cal.n_class, 256, 0, st>>>(ctx->cal.hist, ctx->cal.n_class, ctx->cal.n_bins, d_range);
        CK(cudaMemcpyAsync(range, d_range, sizeof range, cudaMemcpyDeviceToHost, st));
    }
    if (rows) CK(cudaMemcpyAsync(rows, src, sizeof(double) * (size_t)n_rows * ctx->cal.n_bins, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (dptr) *dptr = (void *)src;
    if (cols) { cols[0] = range[1] < 0 ? -1 : range[0]; cols[1] = range[1]; }
    return RMD_OK;
}

static int calibrate_once(rmd_ctx *ctx, int model, const uint8_t *samples, uint64_t seed, int64_t first, int64_t n, int32_t L,
                          const double *class_log, int32_t n_class, double cutoff, double *hist, int32_t n_bins,
                          int64_t *n_ok, int32_t *overflow, float *ms_kernel) {
    if (!ctx || !hist) return fail(ctx, RMD_E_INVALID, "rmd_calibrate: NULL argument");
    if (n <= 0) return RMD_OK;
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    int r;
    if ((r = rmd_calib_open(ctx, model, L, class_log, n_class, cutoff, n_bins))) return r;
    if ((r = rmd_calib_add(ctx, samples, seed, first, n))) return r;
    const size_t nh = (size_t)n_class * n_bins;
    if (ctx->h_calhist_n < nh) {                   // pinned staging: the download is a third of a short call
        if (ctx->h_calhist) cudaFreeHost(ctx->h_calhist);
    if (ctx->h_calcls) cudaFreeHost(ctx->h_calcls);
        ctx->h_calhist = nullptr; ctx->h_calhist_n = 0;
        CK(cudaMallocHost((void **)&ctx->h_calhist, sizeof(double) * nh));
        ctx->h_calhist_n = nh;
    }
    const double *h = ctx->h_calhist;
    if ((r = rmd_calib_read(ctx, ctx->h_calhist, n_ok, overflow, ms_kernel))) return r;
    for (size_t k = 0; k < nh; k++) hist[k] += h[k];
    ctx->cal_open = false;
    return RMD_OK;
}

extern "C" int rmd_calibrate(rmd_ctx *ctx, int model, const uint8_t *samples, int64_t n, int32_t L,
                             const double *class_log, int32_t n_class, double cutoff,
                             double *hist, int32_t n_bins, int64_t *n_ok, int32_t *overflow, float *ms_kernel) {
    if (!samples) return fail(ctx, RMD_E_INVALID, "rmd_calibrate: samples is NULL");
    return calibrate_once(ctx, model, samples, 0, 0, n, L, class_log, n_class, cutoff, hist, n_bins, n_ok, overflow, ms_kernel);
}

extern "C" int rmd_calibrate_synth(rmd_ctx *ctx, int model, uint64_t seed, int64_t first, int64_t n, int32_t L,
                                   const double *class_log, int32_t n_class, double cutoff,
                                   double *hist, int32_t n_bins, int64_t *n_ok, int32_t *overflow, float *ms_kernel) {
    return calibrate_once(ctx, model, nullptr, seed, first, n, L, class_log, n_class, cutoff, hist, n_bins, n_ok, overflow, ms_kernel);
}

extern "C" int rmd_cpt_counts(rmd_ctx *ctx, const uint8_t *letters, int64_t n_rows, int32_t n_cols, int32_t K,
                              const int32_t *tables, int32_t n_tables, const int32_t *run, int32_t n_runs,
                              uint32_t *counts, int32_t *first_row) {
    if (!ctx || !letters || !tables || !run || !counts || !first_row) return fail(ctx, RMD_E_INVALID, "rmd_cpt_counts: NULL argument");
    if (n_rows < 1 || n_rows > 0x7fffffff || n_cols < 1 || K < 2 || K > 16 || n_tables < 1 || n_runs < 1)
        return fail(ctx, RMD_E_INVALID, "rmd_cpt_counts: sizes out of range");
    const long long k3 = (long long)K * K * K, cells = k3 * n_tables;
    if (cells * n_runs > (1LL << 24)) return fail(ctx, RMD_E_LIMIT, "rmd_cpt_counts: %lld runs x tables x cells exceed the limit", cells * n_runs);
    for (int t = 0; t < n_tables; t++)
        for (int c = 0; c < 3; c++) {
            const int v = tables[3 * t + c];
            if (v >= n_cols || (c == 0 && v < 0)) return fail(ctx, RMD_E_INVALID, "rmd_cpt_counts: table %d names column %d", t, v);
        }
    for (int64_t r = 0; r < n_rows; r++) {
        if (run[r] < 0 || run[r] >= n_runs) return fail(ctx, RMD_E_INVALID, "rmd_cpt_counts: row %lld names run %d", (long long)r, run[r]);
        for (int c = 0; c < n_cols; c++)
            if (letters[r * n_cols + c] >= K) return fail(ctx, RMD_E_INVALID, "rmd_cpt_counts: letter code %d at row %lld is not below K", letters[r * n_cols + c], (long long)r);
    }
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    ctx->cal_open = false;                         // shares the scratch buffers
    cudaStream_t st = ctx->stream;
    RESERVE(ctx->d_e0, (size_t)n_rows * n_cols);
    RESERVE(ctx->d_e1, sizeof(int) * 3 * (size_t)n_tables);
    RESERVE(ctx->d_e2, sizeof(int) * (size_t)n_rows);
    RESERVE(ctx->d_e3, sizeof(unsigned int) * (size_t)(cells * n_runs));
    RESERVE(ctx->d_e4, sizeof(int) * (size_t)cells);
    CK(cudaMemcpyAsync(ctx->d_e0.p, letters, (size_t)n_rows * n_cols, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_e1.p, tables, sizeof(int) * 3 * (size_t)n_tables, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_e2.p, run, sizeof(int) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_e3.p, 0, sizeof(unsigned int) * (size_t)(cells * n_runs), st));
    CK(cudaMemsetAsync(ctx->d_e4.p, 0x7f, sizeof(int) * (size_t)cells, st));          // 0x7f7f7f7f: larger than any row
    const long long work = (long long)n_rows * n_tables;
    cpt_count_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>((const uint8_t *)ctx->d_e0.p, n_rows, n_cols, K, (const int *)ctx->d_e1.p,
                                                                    n_tables, (const int *)ctx->d_e2.p, (unsigned int *)ctx->d_e3.p, (int *)ctx->d_e4.p);
    CK(cudaMemcpyAsync(counts, ctx->d_e3.p, sizeof(unsigned int) * (size_t)(cells * n_runs), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(first_row, ctx->d_e4.p, sizeof(int) * (size_t)cells, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (long long c = 0; c < cells; c++) if (first_row[c] == 0x7f7f7f7f) first_row[c] = 0x7fffffff;
    return RMD_OK;
}

extern "C" int rmd_synth_job(rmd_ctx *ctx, int64_t n, int64_t first, uint64_t seq_seed, uint64_t bpp_seed,
                             int32_t win_len, int32_t win_step, double min_bpp_mean, double cutoff,
                             const double *gc_log, const int32_t *gc_class, int64_t *n_win, int64_t *n_bpp, float *ms_generate) {
    if (!ctx || n < 1) return fail(ctx, RMD_E_INVALID, "rmd_synth_job: bad arguments");
    if (ctx->n_models == 0) return fail(ctx, RMD_E_STATE, "rmd_synth_job: call rmd_set_models first");
    if (win_len < 1 || win_len > RMD_FAST_WINDOW || win_step < 1) return fail(ctx, RMD_E_LIMIT, "rmd_synth_job: window length must be 1..%d", RMD_FAST_WINDOW);
    if (!gc_log) return fail(ctx, RMD_E_INVALID, "rmd_synth_job: gc_log is required");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    ctx->have_job = false;
    int64_t seq_off[2] = {0, n};
    int r;
    if ((r = stage_layout(ctx, 1, seq_off, win_len, win_step, -1))) return r;
    if ((r = stage_gc(ctx, gc_log, gc_class))) return r;
    DevJob &J = ctx->job;
    cudaStream_t st = ctx->stream;
    RESERVE(ctx->d_codes2, sizeof(uint32_t) * ((size_t)n / 16 + 2));
    RESERVE(ctx->d_bpp_cnt, sizeof(uint32_t) * (size_t)J.n_win);
    RESERVE(ctx->d_group_off, sizeof(uint32_t) * (size_t)(J.n_win + 1));
    RESERVE(ctx->d_bpp_off, sizeof(long long) * (size_t)(J.n_win + 1));
    J.codes2 = (const uint32_t *)ctx->d_codes2.p; J.codes8 = nullptr; ctx->any_other = false;
    J.min_bpp = min_bpp_mean; J.cutoff = cutoff; J.no_score = !(0.0 >= 0.0 + cutoff);
    job_thresholds(J);
    CK(cudaEventRecord(ctx->ev[0], st));
    const long long words = (n + 15) / 16;
    if (ctx->strand_total > 0 && (first < 0 || first + n > ctx->strand_total))
        return fail(ctx, RMD_E_INVALID, "rmd_synth_job: letters [%lld, %lld) leave the strand of %lld letters", (long long)first, (long long)(first + n), ctx->strand_total);
    synth_codes_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(seq_seed, first, n, ctx->strand_reverse ? ctx->strand_total : 0,
                                                                        (uint32_t *)ctx->d_codes2.p);
    synth_bpp_kernel<false><<<(unsigned)J.n_win, SYN_THREADS, 0, st>>>(J, bpp_seed, (uint32_t *)ctx->d_bpp_cnt.p, nullptr, nullptr, nullptr, nullptr);
    // offsets: 32-bit scan per window, widened on the host (lists stay far below 4 G entries per job)
    int launches = 0;
    if ((r = device_scan(ctx, (const uint32_t *)ctx->d_bpp_cnt.p, J.n_win, (uint32_t *)ctx->d_group_off.p, &launches))) return r;
    std::vector<uint32_t> off32((size_t)J.n_win + 1);
    CK(cudaMemcpyAsync(off32.data(), ctx->d_group_off.p, sizeof(uint32_t) * off32.size(), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    std::vector<long long> off64(off32.begin(), off32.end());
    const long long nb = off64.back();
    ctx->n_bpp = nb;
    RESERVE(ctx->d_bpp_i, sizeof(uint16_t) * (size_t)std::max<long long>(nb, 1));
    RESERVE(ctx->d_bpp_j, sizeof(uint16_t) * (size_t)std::max<long long>(nb, 1));
    RESERVE(ctx->d_bpp_p, sizeof(double) * (size_t)std::max<long long>(nb, 1));
    CK(cudaMemcpyAsync(ctx->d_bpp_off.p, off64.data(), sizeof(long long) * off64.size(), cudaMemcpyHostToDevice, st));
    J.bpp_off = (const long long *)ctx->d_bpp_off.p;
    J.bpp_i = (const uint16_t *)ctx->d_bpp_i.p; J.bpp_j = (const uint16_t *)ctx->d_bpp_j.p; J.bpp_p = (const double *)ctx->d_bpp_p.p;
    J.bpp_ij = nullptr; J.bpp_q = nullptr;
    synth_bpp_kernel<true><<<(unsigned)J.n_win, SYN_THREADS, 0, st>>>(J, bpp_seed, nullptr, J.bpp_off, (uint16_t *)ctx->d_bpp_i.p,
                                                                      (uint16_t *)ctx->d_bpp_j.p, (double *)ctx->d_bpp_p.p);
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (ms_generate) CK(cudaEventElapsedTime(ms_generate, ctx->ev[0], ctx->ev[1]));
    if (n_win) *n_win = J.n_win;
    if (n_bpp) *n_bpp = nb;
    ctx->have_job = true;
    return RMD_OK;
}

extern "C" int rmd_synth_strand(rmd_ctx *ctx, int64_t total, int32_t reverse) {
    if (!ctx || total < 0 || (reverse && total < 1)) return fail(ctx, RMD_E_INVALID, "rmd_synth_strand: bad arguments");
    ctx->strand_total = total;
    ctx->strand_reverse = reverse != 0;
    return RMD_OK;
}

extern "C" int rmd_synth_counts(rmd_ctx *ctx, int64_t n, int64_t first, uint64_t seq_seed, int64_t *counts) {
    if (!ctx || !counts || n < 0 || first < 0) return fail(ctx, RMD_E_INVALID, "rmd_synth_counts: bad arguments");
    if (ctx->strand_total > 0 && first + n > ctx->strand_total)
        return fail(ctx, RMD_E_INVALID, "rmd_synth_counts: letters [%lld, %lld) leave the strand of %lld letters", (long long)first, (long long)(first + n), ctx->strand_total);
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    RESERVE(ctx->d_counts, sizeof(unsigned long long) * RMD_N_CODES);
    CK(cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(unsigned long long) * 4, ctx->stream));
    if (n > 0) {
        const unsigned grid = (unsigned)std::min<long long>((n + 256 * 64 - 1) / (256 * 64), 148 * 16);
        synth_counts_kernel<<<grid, 256, 0, ctx->stream>>>(seq_seed, first, n, ctx->strand_reverse ? ctx->strand_total : 0,
                                                           (unsigned long long *)ctx->d_counts.p);
    }
    CK(cudaMemcpyAsync(counts, ctx->d_counts.p, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return RMD_OK;
}

extern "C" int rmd_download_codes(rmd_ctx *ctx, uint8_t *codes, int64_t n) {
    if (!ctx || !codes) return fail(ctx, RMD_E_INVALID, "rmd_download_codes: NULL argument");
    if (!ctx->have_job) return fail(ctx, RMD_E_STATE, "rmd_download_codes: no resident job");
    if (n > ctx->n_letters) n = ctx->n_letters;
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    if (ctx->any_other) {
        CK(cudaMemcpy(codes, ctx->d_codes8.p, (size_t)n, cudaMemcpyDeviceToHost));
        return RMD_OK;
    }
    std::vector<uint32_t> words((size_t)(n + 15) / 16);
    CK(cudaMemcpy(words.data(), ctx->d_codes2.p, words.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < n; k++) codes[k] = (words[k >> 4] >> (2 * (k & 15))) & 3;
    return RMD_OK;
}

// a job that arrived in the compact transport keeps only that form complete on the device: expand before handing lists back
static int ensure_expanded(rmd_ctx *ctx) {
    if (!ctx->job.bpp_q || ctx->n_bpp <= 0) return RMD_OK;
    expand_bpp_kernel<<<(unsigned)((ctx->n_bpp + 255) / 256), 256, 0, ctx->stream>>>(ctx->job.bpp_ij, ctx->job.bpp_q, ctx->n_bpp, (uint16_t *)ctx->d_bpp_i.p,
                                                                                   (uint16_t *)ctx->d_bpp_j.p, (double *)ctx->d_bpp_p.p);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    ctx->job.bpp_ij = nullptr; ctx->job.bpp_q = nullptr;
    return RMD_OK;
}

extern "C" int rmd_download_bpp(rmd_ctx *ctx, int64_t window, uint16_t *i, uint16_t *j, double *p, int64_t cap, int64_t *n) {
    if (!ctx || !n) return fail(ctx, RMD_E_INVALID, "rmd_download_bpp: NULL argument");
    if (!ctx->have_job) return fail(ctx, RMD_E_STATE, "rmd_download_bpp: no resident job");
    if (window < 0 || window >= ctx->job.n_win) return fail(ctx, RMD_E_INVALID, "rmd_download_bpp: window out of range");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    { int re = ensure_expanded(ctx); if (re) return re; }
    long long off[2];
    CK(cudaMemcpy(off, (const long long *)ctx->d_bpp_off.p + window, sizeof off, cudaMemcpyDeviceToHost));
    const long long cnt = off[1] - off[0];
    *n = cnt;
    const long long take = std::min<long long>(cnt, cap);
    if (take > 0) {
        CK(cudaMemcpy(i, (const uint16_t *)ctx->d_bpp_i.p + off[0], sizeof(uint16_t) * take, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(j, (const uint16_t *)ctx->d_bpp_j.p + off[0], sizeof(uint16_t) * take, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(p, (const double *)ctx->d_bpp_p.p + off[0], sizeof(double) * take, cudaMemcpyDeviceToHost));
    }
    return RMD_OK;
}

extern "C" int rmd_download_job(rmd_ctx *ctx, int64_t *bpp_off, uint16_t *i, uint16_t *j, double *p) {
    if (!ctx || !bpp_off) return fail(ctx, RMD_E_INVALID, "rmd_download_job: NULL argument");
    if (!ctx->have_job) return fail(ctx, RMD_E_STATE, "rmd_download_job: no resident job");
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);
    CK(cudaSetDevice(ctx->device));
    { int re = ensure_expanded(ctx); if (re) return re; }
    CK(cudaMemcpy(bpp_off, ctx->d_bpp_off.p, sizeof(long long) * (ctx->job.n_win + 1), cudaMemcpyDeviceToHost));
    const long long nb = ctx->n_bpp;
    if (i && j && p && nb > 0) {
        CK(cudaMemcpy(i, ctx->d_bpp_i.p, sizeof(uint16_t) * nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(j, ctx->d_bpp_j.p, sizeof(uint16_t) * nb, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(p, ctx->d_bpp_p.p, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    }
    return RMD_OK;
}

#include "rmd_table.cuh"
