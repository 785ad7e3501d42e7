// device_common.cuh — structures and device helpers shared by every kernel of the scan path.
//
// Reference semantics restated here (paths relative to the reference root):
//   window geometry      lib/scanner.py:80-111 (slide)            -> window_geom()
//   letter codes          lib/gcx.py:4-13 (one log-frequency per distinct character)
//   Bayesian-net scoring  lib/node.py:176-277 (NodeSearch.eval)    -> eval_placement()
//
// fp64 throughout, additions issued one by one in the reference's order with explicit
// round-to-nearest intrinsics so the compiler cannot contract or reassociate them.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "../../include/rmdetect_b200.h"

#define FULL_MASK 0xffffffffu

// Per-model data in constant memory: everything that is read warp-uniformly (gate program, sizes,
// table pointers).  Trees and CPTs are read per lane and live in global memory behind the read-only path.
struct CModel {
    int chain_len[2];
    int symmetric, n_tree, n_cfg, n_first, brute, n_leaves;
    const rmd_tree_node *tree;
    const int2 *leaf_dist;
    const double *cpt;
    const int8_t *leaf_offsets; // [n_leaves][chain_len[0] + chain_len[1]]
    const double *ev_table;     // [ev_nclass][ev_nrows]
    const double *ev_thr;       // [ev_nrows]
    int ev_nclass, ev_nrows;
    int trie0, n_trie;          // this model's nodes in c_trie
    int xnode0;                 // this model's first node in the concatenated XNode array
    int cb_slot;                // duplicate bitmap of this model in its group; -1: a trie of at most two nodes needs none
    unsigned short cfg_begin[RMD_MAX_GATE_CFGS + 1];
    rmd_gate_pair pairs[RMD_MAX_GATE_PAIRS];
};
__constant__ CModel c_models[RMD_MAX_MODELS];

// Prefix trie of a model's gate configurations (pair sequences), DFS preorder.  A node stands for "this
// pair is summed", which requires every ancestor pair to be non-zero (the reference's `break`,
// lib/scanner.py:204-206); `mult` configurations run through it, so its value enters the reference's
// sum `mult` times.  Indices (parent, skip) are absolute positions in c_trie.
struct TrieNode {
    int8_t ro, co;            // pair as (row offset from p0, column offset from p1)
    uint8_t mult, model;
    int8_t parent;            // -1: first pair of a configuration
    uint8_t pbit;             // index into c_prel of (parent pair - this pair); 31: no parent
    unsigned short skip;      // first node after this subtree
};
static_assert(sizeof(TrieNode) == 8, "TrieNode layout");
#ifndef TRIE_CAP
#define TRIE_CAP 128          // trie nodes of all resident models (checked by rmd_set_models)
#endif
__constant__ TrieNode c_trie[TRIE_CAP];
__constant__ uint16_t c_prel[32];   // distinct (parent - node) offsets over all tries: (dr & 0xff) | (dc & 0xff) << 8
__constant__ int c_n_prel;
__constant__ unsigned char c_groups[RMD_MAX_MODELS + 1];   // candidate generation handles models [c_groups[g], c_groups[g + 1]) together
__constant__ int c_n_groups;

// The trie nodes of a group again, in the order candidate generation visits them: the nodes of models without SYMMETRIC,
// then those of symmetric models; in each part the root nodes (first pairs of configurations), then the others.
//   x    = (ro + 16) | (co + 16) << 12 | pbit << 24 | symmetric << 29     (see the node loop in scan_kernels.cuh)
//   info = trie node | grandparent << 7 (0x7f: none) | model << 14 | (duplicate bitmap + 1) << 18 |
//          second << 20 (a node other than the first of a trie that has no bitmap) | first node of the model's trie << 21
struct CandNode { uint32_t x, info; };
__constant__ CandNode c_cand[TRIE_CAP];
__constant__ uint32_t c_ancw[TRIE_CAP];                    // per TRIE node: ro | co << 12 | parent << 24 (0x7f: none)
__constant__ unsigned char c_cgrp[RMD_MAX_MODELS + 1][4];  // per group: first candidate index, first node of a SYMMETRIC model, end of the group
#define NODE_SYM (1u << 29)
#define NODE_GUARD 0x800800u

// Search-tree node as the scoring engine reads it (32 bytes, built by rmd_set_models from
// rmd_tree_node; all models concatenated, indices absolute).  The engine keeps the 16 letters from
// p0 on and the 16 letters from p1 on as one 64-bit word, two bits per letter; a letter selector is the
// bit position k = 2 * offset + 32 * chain.  A gap parent is folded into `cpt`, an absent one has stride 0.
struct XNode {
    uint32_t sel;             // k_own | k_parent0 << 8 | k_parent1 << 16 | XSEL_OWN_GAP
    uint32_t cpt;             // index of the node's first CPT entry in the concatenated table
    uint16_t m0, m1;          // CPT strides of the parents' letters
    uint16_t skip, next;      // node after the subtree / next node in preorder; XNODE_END ends the walk
    uint16_t leaf;
    uint8_t model, fb;        // fb = flags | bdepth << 4
    int16_t dlo;              // CHECK_DIST node (lib/node.py:184-192): the walk dies when dlo <= p1 - p0 < dlo + dspan
    uint16_t dspan;           // 0 elsewhere
    uint32_t pad_[2];
};
static_assert(sizeof(XNode) == 32, "XNode layout");
#define XSEL_OWN_GAP (1u << 24)
#define XNODE_MAX_OFFSET 15                  // letters per chain held in the engine's word
#define RMD_SURV_MAX_SEQ 4096                // GC-table rows a survivor record can name
#define XNODE_END 0xffffu
#define LETTER_GAP_SLOT RMD_FAST_WINDOW      // letters[LETTER_GAP_SLOT] holds the gap code (4)

// Prefix screens of the search trees.  The first levels of a top-level subtree read only a handful of letters
// (nine for the sixteen-leaf CL tree down to level 8), so whether ANY node of the screen's last level stays alive
// — the same fp64 sums, the same prune test as the walk itself, lib/node.py:227-248 — is a function of those letters
// and of the sequence's GC table: one bit per letter combination, tabulated per sequence by screen_build_kernel.
// A gate survivor all of whose screens say "dead" cannot reach a leaf and is not walked at all.  The separation
// test (lib/node.py:184-192) is ignored while tabulating, which only errs towards "alive".
#define SCREEN_MAX_LETTERS 9
#define SCREEN_MAX_RUNS 4
#define SCREEN_MAX_PER_MODEL 4
#define SCREEN_CAP (RMD_MAX_MODELS * SCREEN_MAX_PER_MODEL)
struct ScreenDesc {                 // what the scan kernel needs: index = OR of ((letters >> shift) & mask) << dst over the runs
    uint32_t word_off;              // first 32-bit word of this screen's table inside the sequence's block
    uint8_t shift[SCREEN_MAX_RUNS], dst[SCREEN_MAX_RUNS];
    uint32_t mask[SCREEN_MAX_RUNS];
    uint32_t pad_;
};
static_assert(sizeof(ScreenDesc) == 32, "ScreenDesc layout");
struct ScreenBuild {                // what screen_build_kernel needs
    int model, t0, t1, level, npos;
    uint32_t word_off;
    uint8_t pos[SCREEN_MAX_LETTERS + 3];   // letter selectors 2 * offset + 32 * chain, ascending: combination digit k is letter pos[k]
};
__constant__ ScreenBuild c_screen_build[SCREEN_CAP];
__constant__ ScreenDesc c_screens[SCREEN_CAP];

// A resident scan job (device pointers).
struct DevJob {
    int n_seq, n_models, win_len, win_step;
    long long n_win;
    const long long *seq_off;    // [n_seq+1]
    const long long *win_base;   // [n_seq+1] first window index of each sequence
    const uint32_t *codes2;      // 2 bits per letter, 16 letters per word (valid when codes8 == nullptr)
    const uint8_t *codes8;       // one byte per letter, kept only if some letter is not ACGU
    const double *gc_log;        // [n_seq][16]
    const int *gc_class;         // [n_seq][n_models]
    const long long *bpp_off;    // [n_win+1]
    const uint16_t *bpp_i, *bpp_j;
    const double *bpp_p;
    const uint16_t *bpp_ij;      // compact transport kept on the device (rmd_scan): i | j << 8 per entry, values as bpp_q;
    const uint32_t *bpp_q;       // scan_fast_kernel<..., FUSED> then rebuilds bpp_p window by window (nullptr: bpp_i/j/p are complete)
    double min_bpp, cutoff;
    // derived from min_bpp on the host (job_thresholds): as kernel parameters they cost scan_fast_kernel no registers
    double thr_hi, thr_lo;       // min_bpp -+ 1e-12 relative: a fast gate mean outside is decided, inside it is re-summed exactly
    double big_min;              // a mean of values all below this stays below min_bpp
    double inv_sure_unit;        // 1 / (min_bpp (1 + 1e-12)): a value in units of the threshold
    const XNode *xnodes;         // all models' search trees (scoring engine format)
    const double *xcpt;          // all models' CPTs, concatenated
    int n_xnodes, n_xcpt;
    int motif_ok;                // every chain of every model fits the scoring engine's letter word
    int debug;                   // profiling experiments only (RMD_DEBUG): 1 skip scoring, 2 skip the exact gate
    int no_score;                // the root's own prune test fails for this cutoff (lib/node.py:248): no placement can score
    const uint32_t *screen_bits; // prefix-screen tables of the sequences that have them
    const long long *screen_off; // [n_seq] first word of a sequence's block in screen_bits, -1: not tabulated (nullptr: no screens)
    int n_screens;               // screens of all models (c_screens); model m owns [screen0[m], screen0[m + 1])
    unsigned char screen0[RMD_MAX_MODELS + 1];
};

static inline void job_thresholds(DevJob &J) {
    J.thr_hi = J.min_bpp + fabs(J.min_bpp) * 1e-12;
    J.thr_lo = J.min_bpp - fabs(J.min_bpp) * 1e-12;
    J.big_min = J.min_bpp - J.min_bpp * 1e-12;
    J.inv_sure_unit = 1.0 / (J.min_bpp * (1.0 + 1e-12));
}

// Unordered hit record written by the scan kernels; `rank` is its position inside its
// (window, model) group in odometer order, so the ordering pass is a plain scatter.
struct RawHit {
    uint32_t window;     // relative to the scanned range
    uint16_t p0, p1;
    uint16_t model, leaf;
    uint32_t rank;
    double pm, pr;
};
static_assert(sizeof(RawHit) == 32, "RawHit layout");

// per launch pair: window counter of the persistent scan kernel, survivors appended, survivors taken by score_kernel
struct LaunchCtl { unsigned int next_window, pad0; unsigned long long surv_count, head, pad1; };

struct ScanOut {
    RawHit *raw;
    unsigned long long *stats;       // RMD_STATS builds: cycles per phase, summed over warps
    uint4 *surv;                     // gate survivors for score_kernel: window, p0 | p1 << 8 | model << 16 | seq << 20, letters of chain 0, of chain 1
    long long surv_cap;
    struct LaunchCtl *ctl;           // counters of this launch pair (scan_fast_kernel + score_kernel)
    unsigned long long *raw_count;   // [0] records appended, [1] tests_all, [2] tests_pairs
    long long raw_cap;
    uint32_t *gate_pass;             // [n_win_range][n_models]
    uint32_t *group_hits;            // [n_win_range][n_models]
    const uint32_t *other_list;      // windows with a letter other than ACGU, relative to the range (scan_fast_kernel<..., OTHER>)
    const uint32_t *other_count;
    const unsigned int *ready;       // streamed scan in one launch: windows (relative to the range) whose lists have arrived, advanced by the
                                     // copy stream while the kernel runs; nullptr: everything is resident
};

__device__ __forceinline__ int get_code(const DevJob &J, long long g) {
    if (J.codes8) return J.codes8[g];
    return (__ldg(J.codes2 + (g >> 4)) >> (2 * (int)(g & 15))) & 3;
}

// lib/scanner.py:80-111: window k of a sequence starts at k*step, the last one is re-anchored at
// max(0, len - W); in whole-sequence mode (W or step 0, :43-44) the sequence is its only window.
__device__ __forceinline__ void window_geom(const DevJob &J, long long w, int &seq, long long &start, int &wlen) {
    int lo = 0, hi = J.n_seq;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (J.win_base[mid] <= w) lo = mid; else hi = mid;
    }
    seq = lo;
    long long len = J.seq_off[lo + 1] - J.seq_off[lo];
    if (J.win_len == 0 || J.win_step == 0) {
        start = 0;
        wlen = (int)len;
    } else {
        long long k = w - J.win_base[lo], nwin = J.win_base[lo + 1] - J.win_base[lo];
        start = (k == nwin - 1) ? (len > J.win_len ? len - J.win_len : 0) : k * (long long)J.win_step;
        long long rest = len - start;
        wlen = (int)(rest < J.win_len ? rest : J.win_len);
    }
}

// Number of placements the odometer visits (lib/scanner.py:137-157,161-193): p0 = 0 is always
// tried; p0 >= 1 needs p0 + L0 < W; for each p0, p1 starts at 0 (at p0 when SYMMETRIC) and is
// tried once unconditionally, then while p1 + L1 < W.
__device__ __forceinline__ int rows_of(int wlen, int L0) { return wlen - L0 > 1 ? wlen - L0 : 1; }
__device__ __forceinline__ int p1_last(int wlen, int L1, int p1_first) {
    int hi = wlen - L1 - 1;
    return hi > p1_first ? hi : p1_first;
}

// placements the odometer visits in a window of wlen letters, all p0 and p1 (closed form of the loop)
__device__ __forceinline__ unsigned long long odometer_count(const CModel &M, int wlen) {
    const int L1 = M.chain_len[1];
    const int n0 = rows_of(wlen, M.chain_len[0]);
    if (M.symmetric) {                                      // sum over p0 < n0 of max(1, A - p0), A = wlen - L1
        const long long A = wlen - L1, k = A > 0 ? (A < n0 ? A : n0) : 0;
        return (unsigned long long)(k * A - k * (k - 1) / 2 + (n0 - k));
    }
    return (unsigned long long)n0 * (unsigned long long)(p1_last(wlen, L1, 0) + 1);
}

// q / 1e9 for a transport word q < 2^30 (rmd_compact_bpp), correctly rounded, without the division subroutine: a product
// with the rounded reciprocal, its exact residual, one correction.  Equal to the IEEE quotient for every q below 2^30 —
// checked exhaustively on the host with the same three IEEE operations (oracle/rmd_oracle.c: orc_check_div1e9,
// tests/test_ingest.py).
__device__ __forceinline__ double transport_root(uint32_t q) {
    const double x = (double)q, y = __dmul_rn(x, 1e-9);
    return __fma_rn(__fma_rn(-y, 1e9, x), 1e-9, y);
}

struct EvalOut { int leaf; double pm, pr; };

// NodeSearch.eval + the winner of get_solution (lib/node.py:153-277) on the flattened tree.
// LETTER(pos) yields the code (0..15) at a window-relative position.
// `budget` bounds the number of tree nodes visited; *finished tells whether the walk completed
// (an unfinished walk is simply repeated later with no bound — see the two-stage scoring in K1).
template <class LetterFn>
__device__ __forceinline__ EvalOut eval_placement(const CModel &M, LetterFn letter, const double *gc,
                                                  int p0, int p1, double cutoff,
                                                  int budget = 0x7fffffff, bool *finished = nullptr) {
    EvalOut out;
    out.leaf = -1; out.pm = -500.0; out.pr = -500.0;            // lib/node.py:250-251
    if (finished) *finished = true;
    if (!(0.0 >= __dadd_rn(0.0, cutoff))) return out;            // the root's own test (:248)
    double sp[RMD_MAX_BRANCH_DEPTH + 1], sr[RMD_MAX_BRANCH_DEPTH + 1];
    sp[0] = 0.0; sr[0] = 0.0;
    double pm = 0.0, pr = 0.0;
    const uint4 *tree = reinterpret_cast<const uint4 *>(M.tree);
    const int n = M.n_tree;
    int t = 0;
    while (t < n) {
        if (budget-- <= 0) { if (finished) *finished = false; break; }
        const uint4 raw = __ldg(tree + t);
        const int chain = (int)(int8_t)(raw.x & 0xff), oft = (int)(int8_t)((raw.x >> 8) & 0xff);
        const int flags = (raw.y >> 16) & 0xff, bdepth = raw.y >> 24;
        const int skip = raw.z >> 16;
        if (flags & RMD_F_RESTORE) { pm = sp[bdepth - 1]; pr = sr[bdepth - 1]; }
        if (flags & RMD_F_CHECK_DIST) {                           // lib/node.py:184-192
            const int2 dd = __ldg(M.leaf_dist + (raw.z & 0xffff));
            const int d = p1 - p0;
            if ((d >= 0 && d < dd.x) || (d <= 0 && -d < dd.y)) { t = skip; continue; }
        }
        const int nt = oft < 0 ? 4 : letter((chain ? p1 : p0) + oft);   // :199-204
        int row = 0;
        const int npar = (raw.w >> 16) & 0xff;
        if (npar > 0) {                                           // :207-213
            const int pc0 = (int)(int8_t)((raw.x >> 16) & 0xff), po0 = (int)(int8_t)(raw.y & 0xff);
            int c = po0 < 0 ? 4 : letter((pc0 ? p1 : p0) + po0);
            row = c < 5 ? c : 5;
            if (npar > 1) {
                const int pc1 = (int)(int8_t)(raw.x >> 24), po1 = (int)(int8_t)((raw.y >> 8) & 0xff);
                c = po1 < 0 ? 4 : letter((pc1 ? p1 : p0) + po1);
                row = row * RMD_NSYM + (c < 5 ? c : 5);
            }
        }
        double v = __ldg(M.cpt + ((raw.w & 0xffff) + row) * RMD_NSYM + (nt < 5 ? nt : 5));
        const double g = gc[nt];                                  // :223 (slot 4 holds log 0.25)
        if (v == CUDART_INF) v = g;                               // GC-sentinel row (:217-218)
        pm = __dadd_rn(pm, v);                                    // :227-228
        pr = __dadd_rn(pr, g);
        if (!(pm >= __dadd_rn(pr, cutoff))) { t = skip; continue; }   // :248
        if (flags & RMD_F_LEAF) {
            if (pm > out.pm) { out.pm = pm; out.pr = pr; out.leaf = raw.z & 0xffff; }   // :265-275
        } else if (flags & RMD_F_SAVE) {
            sp[bdepth] = pm; sr[bdepth] = pr;
        }
        t++;
    }
    return out;
}

// One screen bit: the walk of tree nodes [t0, t1) (one top-level subtree) restricted to levels <= `level`, on the
// letters `lt` (index = selector / 2: 16 letters of chain 0, then 16 of chain 1).  True when a node of the last level,
// or a shallower leaf, passes its prune test.  Same additions in the same order as eval_placement.
__device__ __forceinline__ bool screen_walk(const CModel &M, const uint8_t *lt, const double *gc, double cutoff,
                                            int t0, int t1, int level) {
    double sp[RMD_MAX_BRANCH_DEPTH + 1], sr[RMD_MAX_BRANCH_DEPTH + 1];
    sp[0] = 0.0; sr[0] = 0.0;
    double pm = 0.0, pr = 0.0;
    const uint4 *tree = reinterpret_cast<const uint4 *>(M.tree);
    int t = t0;
    while (t < t1) {
        const uint4 raw = __ldg(tree + t);
        const int chain = (int)(int8_t)(raw.x & 0xff), oft = (int)(int8_t)((raw.x >> 8) & 0xff);
        const int flags = (raw.y >> 16) & 0xff, bdepth = raw.y >> 24;
        const int skip = raw.z >> 16;
        if (flags & RMD_F_RESTORE) { pm = sp[bdepth - 1]; pr = sr[bdepth - 1]; }
        const int nt = oft < 0 ? 4 : lt[16 * chain + oft];
        int row = 0;
        const int npar = (raw.w >> 16) & 0xff;
        if (npar > 0) {
            const int pc0 = (int)(int8_t)((raw.x >> 16) & 0xff), po0 = (int)(int8_t)(raw.y & 0xff);
            row = po0 < 0 ? 4 : lt[16 * pc0 + po0];
            if (npar > 1) {
                const int pc1 = (int)(int8_t)(raw.x >> 24), po1 = (int)(int8_t)((raw.y >> 8) & 0xff);
                row = row * RMD_NSYM + (po1 < 0 ? 4 : lt[16 * pc1 + po1]);
            }
        }
        double v = __ldg(M.cpt + ((raw.w & 0xffff) + row) * RMD_NSYM + nt);
        const double g = gc[nt];
        if (v == CUDART_INF) v = g;
        pm = __dadd_rn(pm, v);
        pr = __dadd_rn(pr, g);
        if (!(pm >= __dadd_rn(pr, cutoff))) { t = skip; continue; }
        if ((flags & RMD_F_LEAF) || (int)(raw.w >> 24) >= level) return true;
        if (flags & RMD_F_SAVE) { sp[bdepth] = pm; sr[bdepth] = pr; }
        t++;
    }
    return false;
}

// Tabulate the screens of one sequence per blockIdx.z: thread = letter combination, 32 combinations per word.
__global__ void __launch_bounds__(256) screen_build_kernel(const DevJob J, const int *seqs, uint32_t *bits) {
    const ScreenBuild &S = c_screen_build[blockIdx.y];
    const long long combo = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n_combo = 1LL << (2 * S.npos);
    if (combo - (threadIdx.x & 31) >= n_combo) return;           // whole warps leave together
    const int seq = seqs[blockIdx.z];
    const double *gcs = J.gc_log + (long long)seq * RMD_N_CODES;
    double gc[5];
#pragma unroll
    for (int k = 0; k < 5; k++) gc[k] = gcs[k];
    uint8_t lt[32];
#pragma unroll
    for (int k = 0; k < 32; k++) lt[k] = 0;
    bool alive = false;
    if (combo < n_combo) {
        for (int k = 0; k < S.npos; k++) lt[S.pos[k] >> 1] = (uint8_t)((combo >> (2 * k)) & 3);
        alive = screen_walk(c_models[S.model], lt, gc, J.cutoff, S.t0, S.t1, S.level);
    }
    const uint32_t word = __ballot_sync(FULL_MASK, alive);
    if ((threadIdx.x & 31) == 0) bits[J.screen_off[seq] + S.word_off + (combo >> 5)] = word;
}

// exclusive scan of one value per thread over a block of NT threads (NT multiple of 32, <= 1024);
// returns the exclusive prefix and writes the block total to *total.  `tmp` holds NT/32 + 1 words.
template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *tmp, uint32_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(FULL_MASK, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) tmp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < NT / 32 ? tmp[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(FULL_MASK, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < NT / 32) tmp[lane] = winc - w;
        if (lane == 31) tmp[NT / 32] = winc;
    }
    __syncthreads();
    const uint32_t res = tmp[wid] + inc - v;
    *total = tmp[NT / 32];
    __syncthreads();
    return res;
}
