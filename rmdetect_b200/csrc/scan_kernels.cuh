// scan_kernels.cuh — K1: window scan = placement enumeration + bpp gate + Bayesian-net scoring.
//
// Replaces, per window and model (paths relative to the reference root):
//   lib/scanner.py:137-157  the odometer loop of Scanner.search
//   lib/scanner.py:161-193  pointer_inc
//   lib/scanner.py:195-214  __get_bpp_mean  (gate: mean bpp of the MUST pairs >= min_bpp_mean)
//   lib/bpprobs.py:17-18    prob_pair       (sparse symmetric lookup, 0.0 when absent)
//   lib/node.py:176-277     NodeSearch.eval (via eval_placement in device_common.cuh)
//
// Design (one CTA of 128 threads per window, windows up to 160 letters):
//   * the window's bpp coordinate list becomes, in shared memory, a symmetric bit matrix S
//     (candidate generation), an upper-triangle bit matrix with per-word ranks U (value lookup by
//     popcount rank) and the fp64 values in list order;
//   * a placement can pass the gate only if the FIRST MUST pair of some pairing configuration has
//     non-zero probability (every configuration breaks at its first zero, lib/scanner.py:204-206),
//     so candidates are read off S 32 placements at a time: for fixed p0 the first pair (o1, o2)
//     is bit p1 + o2 of row p0 + o1 — one funnel shift per first pair per 32 placements.  The
//     ~70 000 placements of a window cost ~3 000 word operations; only the few percent that are
//     candidates go through the reference's exact gate sum, and only gate survivors are scored;
//   * survivors are compacted through per-warp shared-memory queues (ballot + popc), so the gate
//     and the tree walk run on full warps;
//   * hits are marked in a per-model bit matrix; a popcount prefix over it gives each hit its rank
//     in odometer order, and the hit is then re-scored once to write its record.
#pragma once
#include <math_constants.h>
#include "device_common.cuh"

#define FAST_W RMD_FAST_WINDOW          // 160
#define FAST_WORDS (FAST_W / 32)        // 5
#define S_ROWS (FAST_W + 32)            // rows reachable by p0 + o1 in degenerate short windows
#define S_STRIDE (FAST_WORDS + 1)       // one zero pad word so the funnel shift can read word + 1
#define VAL_CAP 1024                    // bpp values kept in shared memory; beyond that they are read from global
#define SCAN_THREADS 128
#define QCAP 64
#define WORDS_PER_THREAD ((FAST_W * FAST_WORDS + SCAN_THREADS - 1) / SCAN_THREADS)   // 7

struct FastSmem {
    uint32_t S[S_ROWS * S_STRIDE];
    uint2 U[FAST_W * FAST_WORDS];        // .x = bits (j > i), .y = entries before this word in list order
    double vals[VAL_CAP];
    uint32_t H[FAST_W * FAST_WORDS];     // hits of the current model
    double gc[RMD_N_CODES];
    uint32_t q1[SCAN_THREADS / 32][QCAP];
    uint32_t q2[SCAN_THREADS / 32][QCAP];
    uint32_t scan_tmp[SCAN_THREADS / 32 + 1];
    uint32_t gate_pass;
    unsigned long long raw_base;
    uint8_t letters[S_ROWS];
};

// bpp lookup by popcount rank (lib/bpprobs.py:17-18): 0.0 unless the pair is stored
__device__ __forceinline__ double bpp_fast(const FastSmem &sm, const double *vals, int wlen, int b1, int b2) {
    const int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
    if (hi >= wlen || lo == hi) return 0.0;
    const uint2 e = sm.U[lo * FAST_WORDS + (hi >> 5)];
    const uint32_t bit = 1u << (hi & 31);
    if (!(e.x & bit)) return 0.0;
    return vals[e.y + __popc(e.x & (bit - 1))];
}

// __get_bpp_mean (lib/scanner.py:195-214) for one placement per lane.  The loops are warp-uniform
// (constant-memory program), each lane tracks its own `break`.  Returns the gate decision.
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
// non-positive threshold lets every placement through; the kernels then treat every placement as a
// candidate (see `all_candidates` below).

__global__ void __launch_bounds__(SCAN_THREADS)
scan_fast_kernel(const DevJob J, const long long w_begin, const long long w_end, const ScanOut O) {
    __shared__ FastSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long w = w_begin + blockIdx.x;
    if (w >= w_end) return;
    int seq, wlen;
    long long start;
    window_geom(J, w, seq, start, wlen);
    if (wlen <= 0 || wlen > FAST_W) return;                 // long windows: scan_generic_kernel
    const long long gbase = J.seq_off[seq] + start;
    const long long b0 = J.bpp_off[w], b1 = J.bpp_off[w + 1];
    const int nnz = (int)(b1 - b0);
    const double *vals = nnz <= VAL_CAP ? sm.vals : J.bpp_p + b0;

    // ---- stage the window: letters, GC table, zeroed bit matrices -------------------------
    for (int k = tid; k < S_ROWS * S_STRIDE; k += SCAN_THREADS) sm.S[k] = 0;
    for (int k = tid; k < FAST_W * FAST_WORDS; k += SCAN_THREADS) { sm.U[k] = make_uint2(0, 0); sm.H[k] = 0; }
    for (int k = tid; k < S_ROWS; k += SCAN_THREADS) sm.letters[k] = k < wlen ? (uint8_t)get_code(J, gbase + k) : 0;
    if (tid < RMD_N_CODES) sm.gc[tid] = J.gc_log[seq * RMD_N_CODES + tid];
    if (tid == 0) sm.gate_pass = 0;
    __syncthreads();
    for (int e = tid; e < nnz; e += SCAN_THREADS) {
        const int i = J.bpp_i[b0 + e], j = J.bpp_j[b0 + e];
        if (i < j && j < wlen) {
            atomicOr(&sm.S[i * S_STRIDE + (j >> 5)], 1u << (j & 31));
            atomicOr(&sm.S[j * S_STRIDE + (i >> 5)], 1u << (i & 31));
            atomicOr(&sm.U[i * FAST_WORDS + (j >> 5)].x, 1u << (j & 31));
        }
        if (e < VAL_CAP) sm.vals[e] = J.bpp_p[b0 + e];
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
    __syncthreads();

    auto lookup = [&](int a, int b) { return bpp_fast(sm, vals, wlen, a, b); };
    auto letter = [&](int pos) { return (int)sm.letters[pos]; };
    uint32_t *q1 = sm.q1[wid], *q2 = sm.q2[wid];
    unsigned long long tests_all = 0;

    for (int m = 0; m < J.n_models; m++) {
        const CModel &M = c_models[m];
        const int L0 = M.chain_len[0], L1 = M.chain_len[1];
        const int n0 = rows_of(wlen, L0);
        const int p1_max = p1_last(wlen, L1, M.symmetric ? n0 - 1 : 0);
        const int nw = (p1_max >> 5) + 1;
        const int items = n0 * nw;
        int qn1 = 0, qn2 = 0;
        uint32_t npass = 0;

        auto run_eval = [&](int nbatch) {                       // score up to 32 gate survivors
            const bool act = lane < nbatch;
            const uint32_t c = act ? q2[qn2 - nbatch + lane] : 0;
            qn2 -= nbatch;
            __syncwarp();
            if (act) {
                const int p0 = c & 0xffff, p1 = c >> 16;
                const EvalOut r = eval_placement(M, letter, sm.gc, p0, p1, J.cutoff);
                if (r.leaf >= 0) atomicOr(&sm.H[p0 * FAST_WORDS + (p1 >> 5)], 1u << (p1 & 31));
            }
            __syncwarp();
        };
        auto run_gate = [&](int nbatch) {                       // exact gate on up to 32 candidates
            const bool act = lane < nbatch;
            const uint32_t c = act ? q1[qn1 - nbatch + lane] : 0;
            qn1 -= nbatch;
            __syncwarp();
            const bool pass = gate_pass(M, lookup, act, (int)(c & 0xffff), (int)(c >> 16), J.min_bpp);
            const uint32_t pm = __ballot_sync(FULL_MASK, pass);
            if (pass) q2[qn2 + __popc(pm & ((1u << lane) - 1))] = c;
            qn2 += __popc(pm);
            npass += __popc(pm);
            __syncwarp();
            if (qn2 >= 32) run_eval(32);
        };

        for (int base = wid * 32; base < items; base += SCAN_THREADS) {
            const int t = base + lane;
            uint32_t cw = 0;
            int p0 = 0, wd = 0;
            if (t < items) {
                p0 = t / nw; wd = t - p0 * nw;
                if (M.brute || J.min_bpp <= 0.0) cw = 0xffffffffu;      // all_candidates
                else
                    for (int k = 0; k < M.n_first; k++) {       // OR of the first-pair rows, shifted
                        const int col = (wd << 5) + M.first[k].y;
                        const uint32_t *row = &sm.S[(p0 + M.first[k].x) * S_STRIDE + (col >> 5)];
                        cw |= __funnelshift_r(row[0], row[1], col & 31);
                    }
                const int lo = M.symmetric ? p0 : 0;             // odometer range of p1 for this p0
                const int hi = p1_last(wlen, L1, lo);
                const int blo = lo - (wd << 5), bhi = hi - (wd << 5);
                uint32_t keep = bhi < 0 || blo > 31 ? 0u : 0xffffffffu;
                if (blo > 0 && blo <= 31) keep &= 0xffffffffu << blo;
                if (bhi >= 0 && bhi < 31) keep &= 0xffffffffu >> (31 - bhi);
                cw &= keep;
            }
            while (__any_sync(FULL_MASK, cw != 0)) {              // one candidate per lane per round
                const bool has = cw != 0;
                const uint32_t pmask = __ballot_sync(FULL_MASK, has);
                if (has) {
                    const int b = __ffs(cw) - 1;
                    cw &= cw - 1;
                    q1[qn1 + __popc(pmask & ((1u << lane) - 1))] = (uint32_t)p0 | ((uint32_t)((wd << 5) + b) << 16);
                }
                qn1 += __popc(pmask);
                __syncwarp();
                if (qn1 >= 32) run_gate(32);
            }
        }
        if (qn1 > 0) run_gate(qn1);
        if (qn2 > 32) run_eval(32);
        if (qn2 > 0) run_eval(qn2);
        if (lane == 0 && npass) atomicAdd(&sm.gate_pass, npass);
        __syncthreads();

        // ---- rank the hits in odometer order and write their records ------------------------
        uint32_t mine = 0;
        const int k0 = tid * WORDS_PER_THREAD;
#pragma unroll
        for (int k = 0; k < WORDS_PER_THREAD; k++)
            if (k0 + k < FAST_W * FAST_WORDS) mine += __popc(sm.H[k0 + k]);
        uint32_t total, run = block_exclusive_scan<SCAN_THREADS>(mine, sm.scan_tmp, &total);
        const long long grp = (w - w_begin) * J.n_models + m;
        if (tid == 0) {
            O.gate_pass[grp] = sm.gate_pass;
            O.group_hits[grp] = total;
            sm.gate_pass = 0;
            if (total) sm.raw_base = atomicAdd(&O.raw_count[0], (unsigned long long)total);
            // tests_all for this window and model (closed form of the odometer)
            if (M.symmetric) {
                for (int p0 = 0; p0 < n0; p0++) tests_all += (unsigned long long)(p1_last(wlen, L1, p0) - p0 + 1);
            } else {
                tests_all += (unsigned long long)n0 * (unsigned long long)(p1_last(wlen, L1, 0) + 1);
            }
            atomicAdd(&O.raw_count[2], (unsigned long long)O.gate_pass[grp]);
        }
        if (total) {
            __syncthreads();
            const unsigned long long base = sm.raw_base;
            for (int k = 0; k < WORDS_PER_THREAD; k++) {
                if (k0 + k >= FAST_W * FAST_WORDS) break;
                uint32_t hw = sm.H[k0 + k];
                const int p0 = (k0 + k) / FAST_WORDS, wd = (k0 + k) - p0 * FAST_WORDS;
                while (hw) {
                    const int p1 = (wd << 5) + __ffs(hw) - 1;
                    hw &= hw - 1;
                    const unsigned long long slot = base + run++;
                    if ((long long)slot < O.raw_cap) {
                        const EvalOut r = eval_placement(M, letter, sm.gc, p0, p1, J.cutoff);
                        RawHit h;
                        h.window = (uint32_t)(w - w_begin);
                        h.p0 = (uint16_t)p0; h.p1 = (uint16_t)p1;
                        h.model = (uint16_t)m; h.leaf = (uint16_t)r.leaf;
                        h.rank = (uint32_t)(slot - base);
                        h.pm = r.pm; h.pr = r.pr;
                        O.raw[slot] = h;
                    }
                }
                sm.H[k0 + k] = 0;
            }
        }
        __syncthreads();
    }
    if (tid == 0) atomicAdd(&O.raw_count[1], tests_all);
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
scan_generic_kernel(const DevJob J, const long long w_begin, const long long w_end, const ScanOut O,
                    const int take_short) {
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
    if (wlen <= 0 || (wlen <= FAST_W && !take_short)) return;    // short windows: scan_fast_kernel
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
                    h.window = (uint32_t)(w - w_begin);
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
        const long long grp = (w - w_begin) * J.n_models + m;
        O.gate_pass[grp] = s_pass;
        O.group_hits[grp] = group_rank;
        atomicAdd(&O.raw_count[1], tests_all);
        atomicAdd(&O.raw_count[2], (unsigned long long)s_pass);
    }
}
