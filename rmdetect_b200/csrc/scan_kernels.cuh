// scan_kernels.cuh — K1: window scan = placement enumeration + bpp gate + Bayesian-net scoring.
//
// Replaces, per window and model (paths relative to the reference root):
//   lib/scanner.py:137-157  the odometer loop of Scanner.search
//   lib/scanner.py:161-193  pointer_inc
//   lib/scanner.py:195-214  __get_bpp_mean  (gate: mean bpp of the MUST pairs >= min_bpp_mean)
//   lib/bpprobs.py:17-18    prob_pair       (sparse symmetric lookup, 0.0 when absent)
//   lib/node.py:176-277     NodeSearch.eval (tree walk, see also eval_placement in device_common.cuh)
//
// Design (one CTA of 128 threads per window, windows up to 160 letters):
//   * the window's bpp coordinate list becomes, in shared memory, a symmetric bit matrix S
//     ("is this pair non-zero") and an upper-triangle bit matrix with per-word ranks U: the value of
//     pair (i, j) is entry rank + popc(bits below) of the window's list, read through L1;
//   * candidates come from the matrix, not from the 70 000 placements.  The gate mean can reach
//     min_bpp_mean only if one of the summed values is (within 1e-12) at least min_bpp_mean, and a
//     pair is summed only if the pairs before it in its configuration are non-zero (the
//     reference's `break`, lib/scanner.py:204-206).  So every "big" entry (i, j) is matched
//     against every node of the configurations' prefix trie: the node's pair offsets turn (i, j)
//     into a placement (p0, p1), its ancestors are checked in S, and a bitmap removes placements
//     found twice.  Placements never produced this way fail the gate with no work at all;
//   * candidates go through the reference's exact fp64 gate sum, 32 at a time, the program read
//     warp-uniformly from constant memory;
//   * gate survivors are scored by a lane-level engine: every lane carries its own placement and
//     tree-walk state, visits one tree node per warp iteration, and takes the next placement from
//     a per-warp shared-memory queue (ballot + popc compaction) the moment it finishes — lanes
//     never wait for the slowest walk of a batch.  The engine pauses (state kept in registers)
//     when its queue runs dry and resumes later; it is model-agnostic and drains once per window;
//   * a hit is appended to the global record list at once; the ordering pass (post_kernels.cuh)
//     ranks records inside their (window, model) group.
#pragma once
#include <math_constants.h>
#include "device_common.cuh"

#define FAST_W RMD_FAST_WINDOW          // 160
#define FAST_WORDS (FAST_W / 32)        // 5
#define S_ROWS FAST_W                   // windows no longer than a chain go to the generic kernel, so p + o < 160
#define S_STRIDE FAST_WORDS
#define SCAN_THREADS 128
#define SCAN_WARPS (SCAN_THREADS / 32)
#define QCAP 64
#define TRIE_CAP 128                    // trie nodes of all resident models (checked by rmd_set_models)
#define WORDS_PER_THREAD ((FAST_W * FAST_WORDS + SCAN_THREADS - 1) / SCAN_THREADS)   // 7
#define LANE_NONE 0xffffffffu

struct SmModel {                        // per-model fields the scoring engine reads with a per-lane index
    const uint4 *tree;
    const double *cpt;
    const int2 *leaf_dist;
    int n_tree, trie_base;
};

struct FastSmem {
    uint32_t S[S_ROWS * S_STRIDE];
    uint2 U[FAST_W * FAST_WORDS];        // .x = bits (j > i), .y = entries before this word in list order
    double gc[RMD_N_CODES];
    uint32_t CB[2][FAST_W * FAST_WORDS]; // candidates already queued, alternating between models
    SmModel model[RMD_MAX_MODELS];
    uint32_t q1[SCAN_WARPS][QCAP];       // gate candidates (p0 | p1 << 16)
    uint32_t q2[SCAN_WARPS][QCAP];       // gate survivors  (p0 | p1 << 16), model in q2m
    TrieNode trie[TRIE_CAP];
    uint32_t big[SCAN_WARPS][32];        // big entries of the current chunk (i | j << 16)
    uint8_t q2m[SCAN_WARPS][QCAP];
    uint32_t scan_tmp[SCAN_WARPS + 1];
    uint32_t gate_pass[RMD_MAX_MODELS];
    uint8_t letters[S_ROWS];
};

// bpp lookup by popcount rank (lib/bpprobs.py:17-18): 0.0 unless the pair is stored.  The diagonal
// is never stored, so b1 == b2 needs no test of its own.
__device__ __forceinline__ double bpp_fast(const FastSmem &sm, const double *vals, int wlen, int b1, int b2) {
    const int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
    if (hi >= wlen) return 0.0;
    const uint2 e = sm.U[lo * FAST_WORDS + (hi >> 5)];
    const uint32_t bit = 1u << (hi & 31);
    if (!(e.x & bit)) return 0.0;
    return __ldg(vals + e.y + __popc(e.x & (bit - 1)));
}

// __get_bpp_mean (lib/scanner.py:195-214) for one placement per lane: warp-uniform loops over the
// constant-memory program, each lane tracking its own `break`.  The additions and their order are
// exactly the reference's.
template <class Lookup>
__device__ __forceinline__ bool gate_pass(const CModel &M, Lookup lookup, bool active, int p0, int p1,
                                          double min_bpp) {
    double sum = 0.0;
    int cnt = 0;
    for (int c = 0; c < M.n_cfg; c++) {
        bool alive = active;
        const int kend = M.cfg_begin[c + 1];
        for (int k = M.cfg_begin[c]; k < kend; k++) {
            if (!__any_sync(FULL_MASK, alive)) break;
            const rmd_gate_pair pr = M.pairs[k];
            double prob = 0.0;
            if (alive) prob = lookup((pr.c1 ? p1 : p0) + pr.o1, (pr.c2 ? p1 : p0) + pr.o2);
            if (prob == 0.0) alive = false;                 // `break`: this configuration only
            if (alive) { sum = __dadd_rn(sum, prob); cnt++; }
        }
    }
    if (!active) return false;
    const double mean = cnt ? __ddiv_rn(sum, (double)cnt) : 0.0;      // lib/scanner.py:211-214
    return mean >= min_bpp;
}
// With no pair summed the reference's mean is 0.0 and is still compared with min_bpp_mean, so a
// non-positive threshold lets every placement through; every placement is then a candidate.

__global__ void __launch_bounds__(SCAN_THREADS, 9)
scan_fast_kernel(const DevJob J, const long long w_begin, const long long w_end, const long long w_base,
                 const ScanOut O, const int min_wlen) {       // windows [w_begin, w_end); records are relative to w_base
    __shared__ FastSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1;
    const long long w = w_begin + blockIdx.x;
    if (w >= w_end) return;
    int seq, wlen;
    long long start;
    window_geom(J, w, seq, start, wlen);
    if (wlen < min_wlen || wlen > FAST_W) return;           // long or degenerate windows: scan_generic_kernel
    const long long gbase = J.seq_off[seq] + start;
    const long long b0 = J.bpp_off[w], b1 = J.bpp_off[w + 1];
    const int nnz = (int)(b1 - b0);
    const double *vals = J.bpp_p + b0;                        // values stay in global memory (L1): keeping them out
                                                              // of shared memory buys a ninth resident CTA per SM
    const uint32_t win_local = (uint32_t)(w - w_base);

    // ---- stage the window: letters, GC table, bit matrices, model table and tries ------------------
    for (int k = tid; k < S_ROWS * S_STRIDE; k += SCAN_THREADS) sm.S[k] = 0;
    for (int k = tid; k < FAST_W * FAST_WORDS; k += SCAN_THREADS) { sm.U[k] = make_uint2(0, 0); sm.CB[0][k] = 0; }
    for (int k = tid; k < S_ROWS; k += SCAN_THREADS) sm.letters[k] = k < wlen ? (uint8_t)get_code(J, gbase + k) : 0;
    if (tid < RMD_N_CODES) sm.gc[tid] = J.gc_log[seq * RMD_N_CODES + tid];
    if (tid < RMD_MAX_MODELS) sm.gate_pass[tid] = 0;
    {
        int tb = 0;                                           // same layout for every thread
        for (int m = 0; m < J.n_models; m++) {
            const CModel &M = c_models[m];
            for (int k = tid; k < M.n_trie; k += SCAN_THREADS) sm.trie[tb + k] = M.trie[k];
            if (tid == 0) {
                SmModel r;
                r.tree = reinterpret_cast<const uint4 *>(M.tree); r.cpt = M.cpt; r.leaf_dist = M.leaf_dist;
                r.n_tree = M.n_tree; r.trie_base = tb;
                sm.model[m] = r;
            }
            tb += M.n_trie;
        }
    }
    __syncthreads();
    for (int e0 = 0; e0 < nnz; e0 += SCAN_THREADS * 4) {     // four entries per thread in flight
        int ei[4], ej[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int e = e0 + k * SCAN_THREADS + tid;
            ei[k] = ej[k] = 0;
            if (e < nnz) { ei[k] = J.bpp_i[b0 + e]; ej[k] = J.bpp_j[b0 + e]; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int e = e0 + k * SCAN_THREADS + tid, i = ei[k], j = ej[k];
            if (e < nnz && i < j && j < wlen) {
                atomicOr(&sm.S[i * S_STRIDE + (j >> 5)], 1u << (j & 31));
                atomicOr(&sm.S[j * S_STRIDE + (i >> 5)], 1u << (i & 31));
                atomicOr(&sm.U[i * FAST_WORDS + (j >> 5)].x, 1u << (j & 31));
            }
        }
    }
    __syncthreads();
    {   // rank of the first bit of every U word = its position in the (i, j)-sorted list
        uint32_t mine = 0;
        const int k0 = tid * WORDS_PER_THREAD;
#pragma unroll
        for (int k = 0; k < WORDS_PER_THREAD; k++)
            if (k0 + k < FAST_W * FAST_WORDS) mine += __popc(sm.U[k0 + k].x);
        uint32_t total, run = block_exclusive_scan<SCAN_THREADS>(mine, sm.scan_tmp, &total);
#pragma unroll
        for (int k = 0; k < WORDS_PER_THREAD; k++)
            if (k0 + k < FAST_W * FAST_WORDS) { sm.U[k0 + k].y = run; run += __popc(sm.U[k0 + k].x); }
    }
    // a mean of values all below big_min stays below min_bpp: (1 + 2^-53)^256 < 1 + 1e-12
    const bool all_candidates = !(J.min_bpp > 0.0);
    const double big_min = J.min_bpp - J.min_bpp * 1e-12;

    auto lookup = [&](int a, int b) { return bpp_fast(sm, vals, wlen, a, b); };
    uint32_t *q1 = sm.q1[wid], *q2 = sm.q2[wid], *bigl = sm.big[wid];
    uint8_t *q2m = sm.q2m[wid];
    int qn1 = 0, qn2 = 0;                                   // warp-uniform queue fills

    // ---- scoring engine: one tree node per busy lane per iteration (lib/node.py:176-277) ----------
    uint32_t e_c = LANE_NONE;                               // placement in flight (p0 | p1 << 16)
    int e_m = 0, e_t = 0, e_leaf = -1;
    double e_pm = 0.0, e_pr = 0.0, e_bpm = -500.0, e_bpr = -500.0;
    double e_sp[RMD_MAX_BRANCH_DEPTH + 1], e_sr[RMD_MAX_BRANCH_DEPTH + 1];
    const bool root_ok = 0.0 >= __dadd_rn(0.0, J.cutoff);   // the root's own test (:248)
    unsigned long long dbg_iter = 0, dbg_visits = 0;        // RMD_DEBUG & 8: engine statistics

    auto eval_engine = [&](const bool drain) {
        uint32_t bm = __ballot_sync(FULL_MASK, e_c != LANE_NONE);      // busy lanes, kept current below
        for (;;) {
            if (qn2 > 0 && bm != FULL_MASK) {                // hand queued survivors to idle lanes
                const uint32_t im = ~bm;
                const int r = __popc(im & lt_mask);
                const bool take = (im >> lane & 1) && r < qn2;
                if (take) {
                    e_c = q2[qn2 - 1 - r]; e_m = q2m[qn2 - 1 - r];
                    e_t = 0; e_leaf = -1; e_pm = 0.0; e_pr = 0.0; e_bpm = -500.0; e_bpr = -500.0;
                    e_sp[0] = 0.0; e_sr[0] = 0.0;
                }
                qn2 -= min(qn2, __popc(im));
                bm = __ballot_sync(FULL_MASK, e_c != LANE_NONE);
            }
            if (bm == 0) break;
            if (!drain && qn2 == 0 && bm != FULL_MASK) break;   // pause: nothing left to keep every lane fed
            if (J.debug & 8) { dbg_iter++; dbg_visits += __popc(bm); }
            const bool busy = bm >> lane & 1;
            bool done = false;
            if (busy) {
                const SmModel R = sm.model[e_m];
                const int p0 = e_c & 0xffff, p1 = e_c >> 16;
                if (e_t >= R.n_tree || !root_ok) done = true;
                else {
                    const uint4 raw = __ldg(R.tree + e_t);
                    const int chain = (int)(int8_t)(raw.x & 0xff), oft = (int)(int8_t)((raw.x >> 8) & 0xff);
                    const int flags = (raw.y >> 16) & 0xff, bdepth = raw.y >> 24;
                    const int skip = raw.z >> 16;
                    bool dead = false;
                    if (flags & RMD_F_RESTORE) { e_pm = e_sp[bdepth - 1]; e_pr = e_sr[bdepth - 1]; }
                    if (flags & RMD_F_CHECK_DIST) {                           // lib/node.py:184-192
                        const int2 dd = __ldg(R.leaf_dist + (raw.z & 0xffff));
                        const int d = p1 - p0;
                        dead = (d >= 0 && d < dd.x) || (d <= 0 && -d < dd.y);
                    }
                    if (!dead) {
                        const int nt = oft < 0 ? 4 : sm.letters[(chain ? p1 : p0) + oft];     // :199-204
                        int row = 0;
                        const int npar = (raw.w >> 16) & 0xff;
                        if (npar > 0) {                                       // :207-213
                            const int pc0 = (int)(int8_t)((raw.x >> 16) & 0xff), po0 = (int)(int8_t)(raw.y & 0xff);
                            int c = po0 < 0 ? 4 : sm.letters[(pc0 ? p1 : p0) + po0];
                            row = c < 5 ? c : 5;
                            if (npar > 1) {
                                const int pc1 = (int)(int8_t)(raw.x >> 24), po1 = (int)(int8_t)((raw.y >> 8) & 0xff);
                                c = po1 < 0 ? 4 : sm.letters[(pc1 ? p1 : p0) + po1];
                                row = row * RMD_NSYM + (c < 5 ? c : 5);
                            }
                        }
                        double v = __ldg(R.cpt + ((raw.w & 0xffff) + row) * RMD_NSYM + (nt < 5 ? nt : 5));
                        const double g = sm.gc[nt];                           // :223 (slot 4 holds log 0.25)
                        if (v == CUDART_INF) v = g;                           // GC-sentinel row (:217-218)
                        e_pm = __dadd_rn(e_pm, v);                            // :227-228
                        e_pr = __dadd_rn(e_pr, g);
                        dead = !(e_pm >= __dadd_rn(e_pr, J.cutoff));          // :248
                    }
                    if (dead) e_t = skip;
                    else {
                        if (flags & RMD_F_LEAF) {
                            if (e_pm > e_bpm) { e_bpm = e_pm; e_bpr = e_pr; e_leaf = raw.z & 0xffff; }   // :265-275
                        } else if (flags & RMD_F_SAVE) {
                            e_sp[bdepth] = e_pm; e_sr[bdepth] = e_pr;
                        }
                        e_t++;
                    }
                    done = e_t >= R.n_tree;
                }
            }
            const uint32_t dm = __ballot_sync(FULL_MASK, done);
            if (dm == 0) continue;
            const bool hit = done && e_leaf >= 0;
            const uint32_t hm = __ballot_sync(FULL_MASK, hit);
            if (hm) {                                         // append records: one atomic per warp iteration
                unsigned long long base = 0;
                if (lane == __ffs(hm) - 1) base = atomicAdd(&O.raw_count[0], (unsigned long long)__popc(hm));
                base = __shfl_sync(FULL_MASK, base, __ffs(hm) - 1);
                if (hit) {
                    const unsigned long long slot = base + __popc(hm & lt_mask);
                    if ((long long)slot < O.raw_cap) {
                        RawHit h;
                        h.window = win_local;
                        h.p0 = (uint16_t)(e_c & 0xffff); h.p1 = (uint16_t)(e_c >> 16);
                        h.model = (uint16_t)e_m; h.leaf = (uint16_t)e_leaf;
                        h.rank = 0;
                        h.pm = e_bpm; h.pr = e_bpr;
                        O.raw[slot] = h;
                    }
                    atomicAdd(&O.group_hits[(long long)win_local * J.n_models + e_m], 1u);
                }
            }
            if (done) e_c = LANE_NONE;
            bm &= ~dm;
            __syncwarp();
        }
    };

    unsigned long long tests_all = 0;
    for (int m = 0; m < J.n_models; m++) {
        const CModel &M = c_models[m];
        const int L0 = M.chain_len[0], L1 = M.chain_len[1];
        const int n0 = rows_of(wlen, L0);
        uint32_t *CB = sm.CB[m & 1];
        uint32_t npass = 0, n_cand = 0;
        // every warp has finished candidate generation of model m - 1: its bitmap can be cleared for
        // model m + 1 (this model's bitmap was cleared one phase ago, or by the staging code)
        __syncthreads();
        for (int k = tid; k < FAST_W * FAST_WORDS; k += SCAN_THREADS) sm.CB[(m + 1) & 1][k] = 0;

        auto gate_batch = [&](const int nbatch) {               // exact gate on up to 32 candidates
            const bool act = lane < nbatch;
            const uint32_t c = act ? q1[qn1 - nbatch + lane] : 0;
            qn1 -= nbatch;
            __syncwarp();
            const bool pass = gate_pass(M, lookup, act, (int)(c & 0xffff), (int)(c >> 16), J.min_bpp) && !(J.debug & 2);
            const uint32_t pmask = __ballot_sync(FULL_MASK, pass);
            if (pass) { const int at = qn2 + __popc(pmask & lt_mask); q2[at] = c; q2m[at] = (uint8_t)m; }
            qn2 += __popc(pmask);
            npass += __popc(pmask);
            __syncwarp();
            if (J.debug & 1) qn2 = 0;
            if (qn2 >= 32) eval_engine(false);
        };
        // a candidate whose generating value p satisfies p / (pairs in the program) >= min_bpp_mean passes
        // for certain: the fp64 sum is >= p (positive terms, monotone rounding), the count <= that total,
        // and division is monotone — so it skips the exact sum and is queued for scoring directly
        const double n_slots = (double)M.cfg_begin[M.n_cfg];
        auto push_survivors = [&](const bool has, const uint32_t cand) {
            const uint32_t pmask = __ballot_sync(FULL_MASK, has && !(J.debug & 2));
            if (pmask) {
                if (pmask >> lane & 1) { const int at = qn2 + __popc(pmask & lt_mask); q2[at] = cand; q2m[at] = (uint8_t)m; }
                qn2 += __popc(pmask);
                npass += __popc(pmask);
                n_cand += __popc(pmask);
                __syncwarp();
                if (J.debug & 1) qn2 = 0;
                if (qn2 >= 32) eval_engine(false);
            }
        };
        auto push_candidates = [&](const bool has, const uint32_t cand) {
            const uint32_t pmask = __ballot_sync(FULL_MASK, has);
            if (pmask) {
                if (has) q1[qn1 + __popc(pmask & lt_mask)] = cand;
                qn1 += __popc(pmask);
                n_cand += __popc(pmask);
                __syncwarp();
                if (J.debug & 4) qn1 = 0;
                if (qn1 >= 32) gate_batch(32);
            }
        };

        if (all_candidates || M.brute) {
            // every placement of the odometer is a candidate (non-positive threshold, or a model
            // whose MUST pairs sit on one chain and so give no anchor): 32 placements per lane word
            const int p1_max = p1_last(wlen, L1, M.symmetric ? n0 - 1 : 0);
            const int nw = (p1_max >> 5) + 1;
            const int items = n0 * nw;
            for (int base = wid * 32; base < items; base += SCAN_THREADS) {
                const int t = base + lane;
                uint32_t cw = 0;
                int p0 = 0, wd = 0;
                if (t < items) {
                    p0 = t / nw; wd = t - p0 * nw;
                    const int lo = M.symmetric ? p0 : 0;         // odometer range of p1 for this p0
                    const int hi = p1_last(wlen, L1, lo);
                    const int blo = lo - (wd << 5), bhi = hi - (wd << 5);
                    cw = bhi < 0 || blo > 31 ? 0u : 0xffffffffu;
                    if (blo > 0 && blo <= 31) cw &= 0xffffffffu << blo;
                    if (bhi >= 0 && bhi < 31) cw &= 0xffffffffu >> (31 - bhi);
                }
                while (__any_sync(FULL_MASK, cw != 0)) {          // one candidate per lane per round
                    const bool has = cw != 0;
                    const int b = has ? __ffs(cw) - 1 : 0;
                    cw &= cw - 1;
                    push_candidates(has, (uint32_t)p0 | ((uint32_t)((wd << 5) + b) << 16));
                }
            }
        } else {
            // big entries x trie nodes -> placements (see the header of this file)
            const int tb = sm.model[m].trie_base, T = M.n_trie;
            for (int e0 = wid * 32; e0 < nnz; e0 += SCAN_THREADS) {
                const int e = e0 + lane;
                uint32_t ij = 0;
                bool isbig = false;
                if (e < nnz) {
                    const int i = J.bpp_i[b0 + e], j = J.bpp_j[b0 + e];
                    const double p = __ldg(vals + e);
                    isbig = p >= big_min && i < j && j < wlen;
                    // i < 160 and j < 160: bit 15 of the i half flags "passes for certain"
                    ij = (uint32_t)i | ((uint32_t)j << 16) | (__ddiv_rn(p, n_slots) >= J.min_bpp ? 0x8000u : 0u);
                }
                const uint32_t bmask = __ballot_sync(FULL_MASK, isbig);
                if (bmask == 0) continue;
                if (isbig) bigl[__popc(bmask & lt_mask)] = ij;
                __syncwarp();
                const int combos = __popc(bmask) * 2 * T;         // entry x orientation x trie node
                for (int c0 = 0; c0 < combos; c0 += 32) {
                    const int cx = c0 + lane;
                    bool ok = cx < combos, sure = false;
                    uint32_t cand = 0;
                    if (ok) {
                        const int eo = (int)__umulhi((unsigned)cx, M.trie_magic), node = cx - eo * T;
                        const uint32_t en = bigl[eo >> 1];
                        sure = en & 0x8000u;
                        const int row = (eo & 1) ? (int)(en >> 16) : (int)(en & 0x7fff);
                        const int col = (eo & 1) ? (int)(en & 0x7fff) : (int)(en >> 16);
                        TrieNode tn = sm.trie[tb + node];
                        const int p0 = row - tn.ro, p1 = col - tn.co;
                        const int lo = M.symmetric ? p0 : 0;
                        ok = p0 >= 0 && p0 < n0 && p1 >= lo && p1 <= p1_last(wlen, L1, lo);
                        while (ok && tn.parent >= 0) {            // every pair before it must be non-zero
                            tn = sm.trie[tb + tn.parent];
                            const int r = p0 + tn.ro, c = p1 + tn.co;
                            ok = (sm.S[r * S_STRIDE + (c >> 5)] >> (c & 31)) & 1;
                        }
                        if (ok) {
                            const uint32_t bit = 1u << (p1 & 31);
                            ok = !(atomicOr(&CB[p0 * FAST_WORDS + (p1 >> 5)], bit) & bit);   // first time only
                            cand = (uint32_t)p0 | ((uint32_t)p1 << 16);
                        }
                    }
                    push_survivors(ok && sure, cand);
                    push_candidates(ok && !sure, cand);
                }
                __syncwarp();
            }
        }
        if (qn1 > 0) gate_batch(qn1);
        if (lane == 0 && npass) atomicAdd(&sm.gate_pass[m], npass);
        if (lane == 0 && n_cand && J.debug) atomicAdd(&O.raw_count[3], (unsigned long long)n_cand);
        if (tid == 0) {                                         // tests_all: closed form of the odometer
            if (M.symmetric) {                                  // sum over p0 < n0 of max(1, A - p0), A = wlen - L1
                const long long A = wlen - L1, k = A > 0 ? (A < n0 ? A : n0) : 0;
                tests_all += (unsigned long long)(k * A - k * (k - 1) / 2 + (n0 - k));
            } else {
                tests_all += (unsigned long long)n0 * (unsigned long long)(p1_last(wlen, L1, 0) + 1);
            }
        }
    }
    if (J.debug & 1) qn2 = 0;
    eval_engine(true);
    __syncthreads();
    if (tid < J.n_models) {
        O.gate_pass[(long long)win_local * J.n_models + tid] = sm.gate_pass[tid];
        if (sm.gate_pass[tid]) atomicAdd(&O.raw_count[2], (unsigned long long)sm.gate_pass[tid]);
    }
    if (tid == 0) atomicAdd(&O.raw_count[1], tests_all);
    if ((J.debug & 8) && lane == 0) { atomicAdd(&O.raw_count[4], dbg_visits); atomicAdd(&O.raw_count[5], dbg_iter); }
}

// ------------------------------------------------------------------------------------------------
// Generic path for windows longer than FAST_W (large --win-len, or whole-sequence mode,
// lib/scanner.py:43-44).  One CTA per (window, model); placements are visited in odometer order in
// chunks of 256, the bpp list is probed by binary search in global memory.  Exact, not fast.
// ------------------------------------------------------------------------------------------------
#define GEN_THREADS 256

__device__ __forceinline__ double bpp_search(const DevJob &J, long long b0, int nnz, int wlen, int a, int b) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    if (hi >= wlen || lo == hi) return 0.0;
    const uint32_t key = ((uint32_t)lo << 16) | (uint32_t)hi;
    int l = 0, r = nnz;
    while (l < r) {
        const int mid = (l + r) >> 1;
        const uint32_t k = ((uint32_t)J.bpp_i[b0 + mid] << 16) | J.bpp_j[b0 + mid];
        if (k < key) l = mid + 1; else r = mid;
    }
    if (l < nnz && J.bpp_i[b0 + l] == lo && J.bpp_j[b0 + l] == hi) return J.bpp_p[b0 + l];
    return 0.0;
}

__global__ void __launch_bounds__(GEN_THREADS)
scan_generic_kernel(const DevJob J, const long long w_begin, const long long w_end, const long long w_base,
                    const ScanOut O, const int min_fast_wlen) {
    __shared__ uint32_t scan_tmp[GEN_THREADS / 32 + 1];
    __shared__ double gc[RMD_N_CODES];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_pass;
    const int tid = threadIdx.x;
    const long long grp_local = blockIdx.x;
    const int m = (int)(grp_local % J.n_models);
    const long long w = w_begin + grp_local / J.n_models;
    if (w >= w_end) return;
    int seq, wlen;
    long long start;
    window_geom(J, w, seq, start, wlen);
    if (wlen <= 0 || (wlen >= min_fast_wlen && wlen <= FAST_W)) return;    // those windows: scan_fast_kernel
    const CModel &M = c_models[m];
    const long long gbase = J.seq_off[seq] + start;
    const long long b0 = J.bpp_off[w];
    const int nnz = (int)(J.bpp_off[w + 1] - b0);
    if (tid < RMD_N_CODES) gc[tid] = J.gc_log[seq * RMD_N_CODES + tid];
    if (tid == 0) s_pass = 0;
    __syncthreads();
    auto lookup = [&](int a, int b) { return bpp_search(J, b0, nnz, wlen, a, b); };
    auto letter = [&](int pos) { return get_code(J, gbase + pos); };
    const int L0 = M.chain_len[0], L1 = M.chain_len[1];
    const int n0 = rows_of(wlen, L0);
    unsigned long long tests_all = 0;
    uint32_t group_rank = 0, npass = 0;
    for (int p0 = 0; p0 < n0; p0++) {
        const int lo = M.symmetric ? p0 : 0, hi = p1_last(wlen, L1, lo);
        tests_all += (unsigned long long)(hi - lo + 1);
        for (int c0 = lo; c0 <= hi; c0 += GEN_THREADS) {
            const int p1 = c0 + tid;
            const bool act = p1 <= hi;
            const bool pass = gate_pass(M, lookup, act, p0, p1, J.min_bpp);
            EvalOut r; r.leaf = -1;
            if (pass) { npass++; r = eval_placement(M, letter, gc, p0, p1, J.cutoff); }
            uint32_t total;
            const uint32_t rank = block_exclusive_scan<GEN_THREADS>(r.leaf >= 0 ? 1u : 0u, scan_tmp, &total);
            if (total) {
                if (tid == 0) s_base = atomicAdd(&O.raw_count[0], (unsigned long long)total);
                __syncthreads();
                const unsigned long long slot = s_base + rank;
                if (r.leaf >= 0 && (long long)slot < O.raw_cap) {
                    RawHit h;
                    h.window = (uint32_t)(w - w_base);
                    h.p0 = (uint16_t)p0; h.p1 = (uint16_t)p1;
                    h.model = (uint16_t)m; h.leaf = (uint16_t)r.leaf;
                    h.rank = group_rank + rank;
                    h.pm = r.pm; h.pr = r.pr;
                    O.raw[slot] = h;
                }
                group_rank += total;
                __syncthreads();
            }
        }
    }
    if (npass) atomicAdd(&s_pass, npass);
    __syncthreads();
    if (tid == 0) {
        const long long grp = (w - w_base) * J.n_models + m;
        O.gate_pass[grp] = s_pass;
        O.group_hits[grp] = group_rank;
        atomicAdd(&O.raw_count[1], tests_all);
        atomicAdd(&O.raw_count[2], (unsigned long long)s_pass);
    }
}
