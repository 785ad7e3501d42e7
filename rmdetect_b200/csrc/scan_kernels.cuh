// scan_kernels.cuh — K1: window scan = placement enumeration + bpp gate + Bayesian-net scoring.
//
// Replaces, per window and model (paths relative to the reference root):
//   lib/scanner.py:137-157  the odometer loop of Scanner.search
//   lib/scanner.py:161-193  pointer_inc
//   lib/scanner.py:195-214  __get_bpp_mean  (gate: mean bpp of the MUST pairs >= min_bpp_mean)
//   lib/bpprobs.py:17-18    prob_pair       (sparse symmetric lookup, 0.0 when absent)
//   lib/node.py:176-277     NodeSearch.eval (tree walk, see also eval_placement in device_common.cuh)
//
// Design (a team of 128 threads per window, windows up to 160 letters; persistent CTAs of up to 8 teams
// with the models' search trees and CPTs in shared memory):
//   * the window's bpp coordinate list becomes, in shared memory, an upper-triangle bit matrix with
//     per-word ranks (the list position of the entry that opens the word): the value of pair (i, j) is
//     entry rank + popc(bits below) of the window's list, read through L1;
//   * candidates come from the matrix, not from the 70 000 placements.  The gate mean can reach
//     min_bpp_mean only if one of the summed values is (within 1e-12) at least min_bpp_mean, and a
//     pair is summed only if the pairs before it in its configuration are non-zero (the
//     reference's `break`, lib/scanner.py:204-206).  So every "big" entry (i, j), in both
//     orientations, is matched against every node of the configurations' prefix trie: the node's
//     pair offsets turn (i, j) into a placement (p0, p1), its ancestors are checked in the bit
//     matrix, and a bitmap removes placements found twice.  A lane owns one (entry, orientation)
//     and the warp steps through the trie nodes together, so node data are warp-uniform.
//     Placements never produced this way fail the gate with no work at all;
//   * a candidate whose generating value alone lifts the mean over the threshold passes for certain;
//     the others walk the trie lane by lane (depth-first, skipping the subtree of a zero pair) and
//     form the mean from value x multiplicity.  That sum differs from the reference's
//     configuration-by-configuration sum only in rounding (< 1e-13 relative), so only a mean within
//     1e-12 of the threshold is re-summed in the reference's exact order;
//   * gate survivors are scored by a lane-level engine: every lane carries its own placement and
//     tree-walk state, visits one tree node per warp iteration with branch-free code on a
//     pre-decoded 32-byte node, and takes the next placement from a per-warp shared-memory queue
//     the moment it finishes — lanes never wait for the slowest walk of a batch.  The engine
//     pauses (state kept in registers) when its queue runs dry and resumes later; it is
//     model-agnostic and drains once per window;
//   * a hit is appended to the global record list at once; the ordering pass (post_kernels.cuh)
//     ranks records inside their (window, model) group.
// scan_fast_kernel<true> is the fallback for a non-positive threshold or a model whose MUST pairs
// give no anchor (same-chain pair): every placement is a candidate and goes through the exact gate.
#pragma once
#include <math_constants.h>
#include "device_common.cuh"

#define UNROLL_SMALL _Pragma("unroll 1")   // code that runs once per window stays small: 32 warps in different phases share the instruction caches
#define FAST_W RMD_FAST_WINDOW          // 160
#define FAST_WORDS (FAST_W / 32)        // 5
#define UWORDS (FAST_W * FAST_WORDS)    // 800
#ifndef SCAN_THREADS
#define SCAN_THREADS 128
#endif
#define SCAN_WARPS (SCAN_THREADS / 32)
#define Q1_CAP 64                       // gate candidates per warp: < 32 waiting + 32 pushed at once
#define QS_CAP 96                       // gate survivors per warp: < 32 waiting + 32 sure + 32 from one gate batch
#define Q2_CAP 64                       // walks waiting for the scoring engine: < 32 waiting + 32 from one screen batch
#define WL_CAP 48                       // < 16 waiting + 32 from one chunk of entries
#define S2_CAP 96                       // < 32 waiting + 2 x 32 pushed by one step of the node loop
#ifndef CB_SLOTS
#define CB_SLOTS 1                      // duplicate bitmaps resident together (models whose gate trie has more than two nodes)
#endif
#ifndef SCAN_MAX_THREADS
#define SCAN_MAX_THREADS 1024           // a CTA holds up to 8 teams of SCAN_THREADS
#endif
#ifndef REFILL_MIN_IDLE
#define REFILL_MIN_IDLE 5               // the scoring engine hands out queued work once this many lanes are idle
#endif
#ifndef TAIL_ENTRIES
#define TAIL_ENTRIES 192                // the last entries of a window are handed out in smaller rows
#define TAIL_ROW 8
#endif
#define STK_SLOTS 3                     // save slots of the scoring engine kept in shared memory (per lane)
#define SCAN_CTA_FIXED (TRIE_CAP * 32)   // bytes of CTA-wide state in front of the teams' blocks: gate tries (8 B per node), cinfo, ancw, cnw (16 B per node)
#define ERR_BAD_LIST 1ull               // O.raw_count[7]: a window's list is not a strictly increasing upper triangle
#define ERR_BAD_VALUE 2ull              // a stored probability is negative or not a number
#define ERR_TIMEOUT 4ull                // a streamed scan saw no list arrive for STREAM_STALL_NS
#define STREAM_STALL_NS 200000000ull    // 0.2 s

struct WarpQueues {                      // one warp's work lists (one base address per warp)
    uint32_t wl[WL_CAP];                 // big entries waiting for a full row, i | j << 8 | sure-mask << 16
    uint32_t s2[S2_CAP];                 // (placement, trie node) pairs that passed the parent test
    uint32_t q1[Q1_CAP];                 // gate candidates  p0 | p1 << 8 | model << 16
    uint32_t qs[QS_CAP];                 // gate survivors, same format: on their way through the prefix screens
    uint32_t q2[Q2_CAP];                 // survivors some screen calls alive: the scoring engine's queue
};

template <int WCAP>
struct __align__(16) FastSmemT {
    static constexpr int WORDS = WCAP / 32, UW = WCAP * (WCAP / 32);
    uint32_t Ub[UW];                 // bit j of word (i, j >> 5): pair (i, j), i < j, is stored
    uint16_t Ur[UW];                 // entries before this word in list order
    uint32_t CB[CB_SLOTS][UW];       // placements already queued, per bitmap-using model of the current group
    WarpQueues wq[SCAN_WARPS];
    double2 stk[STK_SLOTS][SCAN_THREADS]; // running sums saved at gap branches, [slot][thread]: conflict-free 16-byte accesses
    // per lane, right behind the save slots and with their stride, so that one address register serves both: random-model sum
    // of the walk's best leaf so far (read when the walk ends), window of the walk in flight, its best leaf | model << 16
    // (-1: none yet)
    struct LaneRec { double best_pr; uint32_t walk_win; int walk_leaf; } rec[SCAN_THREADS];
    const uint32_t *tbl;                  // prefix-screen tables of the current sequence (nullptr: every survivor is walked)
    int cur_seq;                          // sequence whose GC table is in gc
    uint32_t win_local;                   // window being generated (records are relative to w_base)
    int nnz_s;                            // entries of its list (read where a register would be live across the whole window)
    double2 zero2;                       // (0, 0)
    // per trie node in candidate order: the odometer's ranges in this window (lib/scanner.py:161-193) as
    // (n0 - 1 + 0x800) | (hi + 0x800) << 12, next to the static words of the node in cnw (see the node loop)
    uint32_t nodey[TRIE_CAP];
    double gc[RMD_N_CODES];
    uint32_t bounds[RMD_MAX_MODELS];     // n0 | p1 upper bound << 8 | symmetric << 16
    uint32_t gate_pass[RMD_MAX_MODELS];
    uint16_t xnode0[RMD_MAX_MODELS], trie0[RMD_MAX_MODELS], trie_end[RMD_MAX_MODELS];
    int8_t cb_slot[RMD_MAX_MODELS];
    uint8_t screen0[RMD_MAX_MODELS + 1]; // model m owns screens [screen0[m], screen0[m + 1])
    uint32_t chunk_ctr[2];               // next entry chunk of the window, per model group (alternating)
    uint32_t next[2];                    // window taken from the global counter (alternating slots)
    uint16_t prel[32];                   // distinct parent-relative offsets of the tries: (dr & 0xff) | dc << 8
    uint8_t letters[WCAP + 16];
    uint32_t l2[WCAP / 16 + 2];        // the letters again, two bits each (a letter other than ACGU: its code's low bits)
    uint32_t lo[WCAP / 32 + 3];        // one bit per letter: it is not one of ACGU (OTHER variant)
    uint32_t plain;                      // every letter of the window is one of ACGU
};
typedef FastSmemT<FAST_W> FastSmem;       // the common case: windows up to 160 letters (FastSmemT<256>: up to 256, five teams per CTA)

// is pair (b1, b2) stored?  (lib/bpprobs.py:17-18; the diagonal is never stored)
template <class SM>
__device__ __forceinline__ bool bpp_present(const SM &sm, int wlen, int b1, int b2) {
    const int lo = min(b1, b2), hi = max(b1, b2);
    return lo >= 0 && hi < wlen && ((sm.Ub[lo * SM::WORDS + (hi >> 5)] >> (hi & 31)) & 1u);
}

// the same for a pair inside the window given as fields b1 | b2 << 12 (the diagonal is never stored)
template <class SM>
__device__ __forceinline__ bool bpp_present_f(const SM &sm, uint32_t f) {
    const uint32_t b1 = f & 0xfffu, b2 = f >> 12;
    const uint32_t lo = min(b1, b2), hi = max(b1, b2);
    return (sm.Ub[lo * SM::WORDS + (hi >> 5)] >> (hi & 31)) & 1u;
}

// bpp lookup by popcount rank: 0.0 unless the pair is stored
template <bool COHERENT = false, class SM>
__device__ __forceinline__ double bpp_fast(const SM &sm, const double *vals, int wlen, int b1, int b2) {
    const int lo = min(b1, b2), hi = max(b1, b2);
    if (hi >= wlen) return 0.0;
    const int k = lo * SM::WORDS + (hi >> 5);
    const uint32_t w = sm.Ub[k], bit = 1u << (hi & 31);
    if (!(w & bit)) return 0.0;
    const double *at = vals + sm.Ur[k] + __popc(w & (bit - 1));
    return COHERENT ? *at : __ldg(at);                  // coherent: the team itself wrote the value (FUSED)
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

// tests_all of one window over all models; out of line: one thread per window calls it, and the scan kernel's loop body has
// to stay small (instruction caches)
__device__ __noinline__ unsigned long long odometer_count_all(const int n_models, const int wlen) {
    unsigned long long tests_all = 0;
    for (int m = 0; m < n_models; m++) tests_all += odometer_count(c_models[m], wlen);
    return tests_all;
}

// barrier of one team (4 warps); barrier 0 stays with __syncthreads()
__device__ __forceinline__ void team_sync(const int team) {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(SCAN_THREADS) : "memory");
}

// Persistent kernel: one CTA per SM slot, `blockDim.x / 128` teams of four warps; a team takes the next
// window from a global counter, so a slow window delays nobody else.  The models' search trees and CPTs
// sit in shared memory once per CTA (TABLES: they fit), next to the teams' window state.
// LSTK: the model set branches deeper than STK_SLOTS, the engine's save slots go to local memory.
// FUSED: the job's lists are on the device in the compact transport only (rmd_scan); every team rebuilds the values of
// its own window while staging it and reads them back with coherent loads.
// OTHER: the launch takes the windows named by O.other_list — those that hold a letter other than ACGU (N runs, IUPAC
// codes), which the plain launch skips.  Such a letter has its own log-frequency (lib/gcx.py:4-13) and no CPT entry
// (lib/node.py:215-223: unknown_prob): the engine carries one more bit per letter, reads the letter's code from the
// window when the bit is set, and finishes its walks before the team leaves the window; placements that touch such a
// letter bypass the prefix screens, whose tables are indexed by ACGU.
template <bool BRUTE, bool TABLES, bool LSTK, bool FUSED = false, bool OTHER = false, int WCAP = FAST_W>
__global__ void __launch_bounds__(SCAN_MAX_THREADS, 1)
scan_fast_kernel(const DevJob J, const long long w_begin, const long long w_end, const long long w_base,
                 const ScanOut O, const int min_wlen, unsigned int *next_window) {       // windows [w_begin, w_end); records are relative to w_base
    typedef FastSmemT<WCAP> Smem;                            // WCAP: the longest window this instantiation stages (160 or 256)
    extern __shared__ uint4 dyn_smem[];
    int tid = threadIdx.x % SCAN_THREADS, lane = tid & 31;
    asm volatile("mov.u32 %0, %0;" : "+r"(tid));          // opaque like the team offset below: no re-derivation from %tid
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));
    const int wid = tid >> 5, team = threadIdx.x / SCAN_THREADS;
    uint32_t lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const int n_models = J.n_models;
    // ---- shared memory: the teams' state first (a team's base is team * sizeof(Smem)), then the CTA-wide tables.
    // The team's byte offset goes through an opaque move: left to itself the compiler re-derives it from
    // %tid and the kernel parameters in front of most shared-memory accesses (13 % of all instructions).
    // The gate tries (static, read by every team) sit once per CTA at the very start: a compile-time address.
    const TrieNode *const strie = reinterpret_cast<const TrieNode *>(dyn_smem);
    const uint32_t *const cinfo = reinterpret_cast<const uint32_t *>(dyn_smem) + TRIE_CAP * 2;   // CandNode::info, candidate order
    const uint32_t *const ancw = cinfo + TRIE_CAP;                                               // c_ancw, trie order
    // static words of a node in candidate order, as the node loop reads them (one 16-byte broadcast load per node):
    //   x = (ro + 16) | (co + 16) << 12, y = 1 << (index of the parent-relative pair; 31: no parent),
    //   z = node << 24 (its place in a queued pair), w = need: a generating value of at least need x min_bpp passes the
    //   gate for certain (255: never)
    const uint4 *const cnw = reinterpret_cast<const uint4 *>(dyn_smem) + TRIE_CAP;
    uint32_t sm_off = (uint32_t)SCAN_CTA_FIXED + (uint32_t)team * (uint32_t)sizeof(Smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(sm_off));
    Smem &sm = *reinterpret_cast<Smem *>(reinterpret_cast<char *>(dyn_smem) + sm_off);
    const int t_words = SCAN_CTA_FIXED / 16 + (int)(blockDim.x / SCAN_THREADS) * (int)(sizeof(Smem) / 16);   // in uint4
    const int x_words = TABLES ? 2 * J.n_xnodes : 0, c_words = TABLES ? (J.n_xcpt + 1) / 2 : 0;
    const int s_words = 2 * J.n_screens;
    uint4 *sx = dyn_smem + t_words;
    double *scpt = reinterpret_cast<double *>(dyn_smem + t_words + x_words);
    const ScreenDesc *const scr = reinterpret_cast<const ScreenDesc *>(dyn_smem + t_words + x_words + c_words);
    for (int k = threadIdx.x; k < s_words; k += blockDim.x) dyn_smem[t_words + x_words + c_words + k] = reinterpret_cast<const uint4 *>(c_screens)[k];
    if (TABLES) {
        const uint4 *gx = reinterpret_cast<const uint4 *>(J.xnodes);
        for (int k = threadIdx.x; k < x_words; k += blockDim.x) sx[k] = __ldg(gx + k);
        for (int k = threadIdx.x; k < J.n_xcpt; k += blockDim.x) scpt[k] = __ldg(J.xcpt + k);
    }
    if (tid < n_models) {
        const CModel &M = c_models[tid];
        sm.xnode0[tid] = (uint16_t)M.xnode0;
        sm.trie0[tid] = (uint16_t)M.trie0;
        sm.trie_end[tid] = (uint16_t)(M.trie0 + M.n_trie);
        sm.cb_slot[tid] = (int8_t)M.cb_slot;
    }
    if (tid <= RMD_MAX_MODELS) sm.screen0[tid] = J.screen0[tid];
    if (tid < 32) sm.prel[tid] = c_prel[tid];
    if (tid == 0) sm.zero2 = make_double2(0.0, 0.0);
    {
        // Sure passes.  A node's value enters the reference's sum once per configuration that runs through the node
        // (`mult` times), and the count never exceeds the model's MUST slots: mean = sum / count >= mult * p / slots
        // (positive terms, monotone rounding, division monotone).  So p >= need * min_bpp * (1 + 1e-12) with
        // need = ceil(slots / mult) passes for certain; the factor covers the roundings.  255: never.
        const int nt = c_models[n_models - 1].trie0 + c_models[n_models - 1].n_trie;
        for (int k = tid; k < nt; k += SCAN_THREADS) {
            const CandNode cn = c_cand[k];
            const TrieNode tn = c_trie[cn.info & 0x7fu];
            if (team == 0) {
                const_cast<TrieNode *>(strie)[k] = c_trie[k];
                const_cast<uint32_t *>(cinfo)[k] = cn.info;
                const_cast<uint32_t *>(ancw)[k] = c_ancw[k];
            }
            const int slots = c_models[tn.model].cfg_begin[c_models[tn.model].n_cfg];
            const uint32_t need = (uint32_t)min((slots + tn.mult - 1) / max((int)tn.mult, 1), 255);
            if (team == 0) const_cast<uint4 *>(cnw)[k] = make_uint4(cn.x & 0xffffffu, 1u << ((cn.x >> 24) & 31u), (uint32_t)k << 24, need);
        }
    }
    __syncthreads();
    const uint4 *xn = TABLES ? sx : reinterpret_cast<const uint4 *>(J.xnodes);
    const double *xcpt = TABLES ? scpt : J.xcpt;
    // the root's own test (lib/node.py:248: 0.0 >= 0.0 + cutoff) is taken on the host, once per job: when it fails nothing can score
#ifdef RMD_TUNING
    const bool no_score = J.no_score || (J.debug & 1);     // profiling experiment: gate only
#else
    const bool no_score = J.no_score != 0;
#endif
    uint32_t wq_off = sm_off + (uint32_t)offsetof(Smem, wq) + (uint32_t)wid * (uint32_t)sizeof(WarpQueues);
    asm volatile("mov.u32 %0, %0;" : "+r"(wq_off));
    WarpQueues &wq = *reinterpret_cast<WarpQueues *>(reinterpret_cast<char *>(dyn_smem) + wq_off);
    uint32_t *const q1 = wq.q1, *const q2 = wq.q2, *const qs = wq.qs;

    int qn1 = 0, qn2 = 0, qns = 0;                          // warp-uniform queue fills
    if (tid == 0) { sm.cur_seq = -1; sm.tbl = nullptr; }    // team-wide state lives in shared memory: registers are the scarce resource

    // ---- scoring engine: one tree node per busy lane per iteration (lib/node.py:176-277) ------
    // A walk needs nothing of its window but the chains' letters (a register pair) and the sequence's GC table,
    // so walks in flight carry over into the team's next window: the engine drains only when the sequence
    // changes and at the end of the kernel.
    uint32_t e_c = 0, e_t = XNODE_END;                  // placement p0 | p1 << 16; node (XNODE_END: lane idle)
    uint32_t e_o = 0;                                   // OTHER: bit k: letter k from p0 on is not ACGU, bit 16 + k: from p1 on
    unsigned long long e_m = 0;                         // 16 letters from p0 on | 16 letters from p1 on << 32, two bits each
    double e_pm = 0.0, e_pr = 0.0, e_bpm = -500.0;      // the best leaf's other sum and the walk's window wait in shared memory
    // sums saved at gap branches: slot k - 1 holds the sums of the branch node at (renumbered) depth k; a node that
    // restores at depth 0 is a child of the implicit root and restarts from zero
    double2 e_stk[LSTK ? RMD_MAX_BRANCH_DEPTH + 2 : 1];
    e_stk[0] = make_double2(0.0, 0.0);
    double2 *const s_stk = &sm.stk[0][tid];
    typename Smem::LaneRec *const s_rec = reinterpret_cast<typename Smem::LaneRec *>(s_stk + STK_SLOTS * SCAN_THREADS);   // = &sm.rec[tid]
    static_assert(sizeof(typename Smem::LaneRec) == sizeof(double2) && offsetof(Smem, rec) == offsetof(Smem, stk) + sizeof(double2) * STK_SLOTS * SCAN_THREADS,
                  "the per-lane records follow the save slots");
    int *const s_leaf = &s_rec->walk_leaf;
    double *const s_bpr = &s_rec->best_pr;
    uint32_t *const s_win = &s_rec->walk_win;
#ifdef RMD_STATS
    unsigned long long dbg_iter = 0, dbg_visits = 0;    // engine statistics
#endif

    auto engine = [&](const bool drain) {
        uint32_t bm = __ballot_sync(FULL_MASK, e_t != XNODE_END);     // busy lanes, kept current below
        for (;;) {
            if (qn2 > 0 && __popc(~bm) >= (drain ? 1 : REFILL_MIN_IDLE)) {   // hand queued survivors to idle lanes
                const uint32_t im = ~bm;
                const int r = __popc(im & lt_mask);
                if ((im >> lane & 1) && r < qn2) {
                    const uint32_t c = q2[qn2 - 1 - r];
                    const uint32_t p0 = c & 0xffu, p1 = (c >> 8) & 0xffu;
                    e_c = p0 | (p1 << 16);
                    *s_win = sm.win_local;
                    e_t = sm.xnode0[c >> 16];
                    e_m = (unsigned long long)__funnelshift_r(sm.l2[p0 >> 4], sm.l2[(p0 >> 4) + 1], (p0 & 15u) * 2) |
                          ((unsigned long long)__funnelshift_r(sm.l2[p1 >> 4], sm.l2[(p1 >> 4) + 1], (p1 & 15u) * 2) << 32);
                    if (OTHER) e_o = (__funnelshift_r(sm.lo[p0 >> 5], sm.lo[(p0 >> 5) + 1], p0 & 31u) & 0xffffu) |
                                     (__funnelshift_r(sm.lo[p1 >> 5], sm.lo[(p1 >> 5) + 1], p1 & 31u) << 16);
                    *s_leaf = -1; e_pm = 0.0; e_pr = 0.0; e_bpm = -500.0;    // lib/node.py:250-251 (a hit always stores its own best_pr)
                }
                qn2 -= min(qn2, __popc(im));
                bm = __ballot_sync(FULL_MASK, e_t != XNODE_END);
            }
            if (bm == 0) break;
            if (!drain && qn2 == 0) break;                   // pause: the walks in flight wait for company
#ifdef RMD_STATS
            dbg_iter++; dbg_visits += __popc(bm);
#endif
            const bool busy = e_t != XNODE_END;
            if (busy) {
                const uint4 a = xn[2 * e_t];                      // sel, cpt, m0 | m1 << 16, skip | next << 16
                const uint2 b = *reinterpret_cast<const uint2 *>(xn + 2 * e_t + 1);   // leaf | model << 16 | flags << 24 | bdepth << 28, dlo | dspan << 16
                int nt = (a.x & XSEL_OWN_GAP) ? 4 : (int)((uint32_t)(e_m >> (a.x & 63u)) & 3u);          // :199-213
                uint32_t c0 = (uint32_t)(e_m >> ((a.x >> 8) & 63u)) & 3u, c1 = (uint32_t)(e_m >> ((a.x >> 16) & 63u)) & 3u;
                int gslot = nt;
                if (OTHER && e_o) {                                // a letter other than ACGU: CPT column / row 5 (unknown_prob), its own GC slot
                    if (!(a.x & XSEL_OWN_GAP) && ((e_o >> ((a.x & 63u) >> 1)) & 1u)) {
                        const uint32_t p = (a.x & 32u) ? (e_c >> 16) : (e_c & 0xffffu);
                        gslot = sm.letters[p + ((a.x & 31u) >> 1)];
                        nt = 5;
                    }
                    if ((e_o >> (((a.x >> 8) & 63u) >> 1)) & 1u) c0 = 5;      // an absent parent has stride 0
                    if ((e_o >> (((a.x >> 16) & 63u) >> 1)) & 1u) c1 = 5;
                }
                double v = xcpt[a.y + nt + c0 * (a.z & 0xffffu) + c1 * (a.z >> 16)];
                const double g = sm.gc[gslot];                    // :223 (slot 4 holds log 0.25)
                if (v == CUDART_INF) v = g;                       // GC-sentinel row (:217-218)
                const uint32_t fl = b.x;
                const int bdepth = fl >> 28;
                if (fl & (RMD_F_RESTORE << 24)) {
                    const double2 s = LSTK ? e_stk[bdepth] : *(bdepth ? s_stk + (bdepth - 1) * SCAN_THREADS : &sm.zero2);
                    e_pm = s.x; e_pr = s.y;
                }
                // :184-192 as one interval test: chains closer than the configuration allows (empty unless CHECK_DIST)
                const int d = (int)(e_c >> 16) - (int)(e_c & 0xffffu);
                const bool apart = (uint32_t)(d - (int)(short)(b.y & 0xffffu)) >= (b.y >> 16);
                e_pm = __dadd_rn(e_pm, v);                        // :227-228
                e_pr = __dadd_rn(e_pr, g);
                // a node that fails is left through its skip link, and whatever comes next restores the sums
                const bool alive = apart && e_pm >= __dadd_rn(e_pr, J.cutoff);       // :248
                if (alive && (fl & (RMD_F_LEAF << 24)) && e_pm > e_bpm) {            // :265-275
                    e_bpm = e_pm; *s_bpr = e_pr; *s_leaf = (int)(fl & 0xffffffu);
                }
                if (alive && (fl & (RMD_F_SAVE << 24))) {
                    if (LSTK) e_stk[bdepth + 1] = make_double2(e_pm, e_pr);
                    else s_stk[bdepth * SCAN_THREADS] = make_double2(e_pm, e_pr);
                }
                e_t = alive ? (a.w >> 16) : (a.w & 0xffffu);
            }
            const uint32_t dm = __ballot_sync(FULL_MASK, busy && e_t == XNODE_END);
            if (dm == 0) continue;
            const int e_leaf = *s_leaf;
            const bool hit = busy && e_t == XNODE_END && e_leaf >= 0;
            const uint32_t hm = __ballot_sync(FULL_MASK, hit);
            if (hm) {                                     // append records: one atomic per warp iteration
                unsigned long long base = 0;
                if (lane == __ffs(hm) - 1) base = atomicAdd(&O.raw_count[0], (unsigned long long)__popc(hm));
                base = __shfl_sync(FULL_MASK, base, __ffs(hm) - 1);
                if (hit) {
                    const unsigned long long slot = base + __popc(hm & lt_mask);
                    const int model = e_leaf >> 16;
                    if ((long long)slot < O.raw_cap) {
                        RawHit h;
                        h.window = *s_win;
                        h.p0 = (uint16_t)(e_c & 0xffffu); h.p1 = (uint16_t)(e_c >> 16);
                        h.model = (uint16_t)model; h.leaf = (uint16_t)(e_leaf & 0xffff);
                        h.rank = 0;
                        h.pm = e_bpm; h.pr = *s_bpr;
                        O.raw[slot] = h;
                    }
                    atomicAdd(&O.group_hits[(long long)*s_win * n_models + model], 1u);
                }
            }
            bm &= ~dm;
        }
    };

    for (int turn = 0;; turn ^= 1) {
        if (tid == 0) {
            unsigned int nx = atomicAdd(next_window, 1u);
            if (FUSED && O.ready != nullptr && w_begin + nx < w_end) {
                // the lists of this window may still be on their way (rmd_scan copies them while this one launch runs):
                // wait for the copy stream's progress mark, which it writes after the data
                unsigned int have;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(O.ready) : "memory");
                if (have <= nx) {
                    // pieces arrive every quarter of a millisecond; a mark that stands still for STREAM_STALL_NS means that the
                    // copies cannot run beside this launch (a profiler or debugger that serialises the GPU's work): give up,
                    // the host scans again once the lists are all there
                    unsigned long long t0, t1;
                    unsigned int seen = have;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                    do {
                        __nanosleep(400);
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(O.ready) : "memory");
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (have != seen) { seen = have; t0 = t1; }
                        if (t1 - t0 > STREAM_STALL_NS) { atomicOr(&O.raw_count[7], ERR_TIMEOUT); nx = 0xffffffffu; break; }
                    } while (have <= nx);
                }
            }
            sm.next[turn] = nx;
        }
        team_sync(team);
        const bool last = OTHER ? sm.next[turn] >= *O.other_count : w_begin + sm.next[turn] >= w_end;
        const long long w = w_begin + (OTHER ? (last ? w_end - w_begin : (long long)O.other_list[sm.next[turn]]) : (long long)sm.next[turn]);
        const int cur_seq = sm.cur_seq;
        int seq = cur_seq, wlen = 0;
        long long start = 0;
        if (!last) window_geom(J, w, seq, start, wlen);
        const bool new_seq = seq != cur_seq;
        if (last || new_seq) {                                // walks in flight read the old sequence's GC table: finish them
            qn2 = 0;
            engine(true);
            if (last) break;
            team_sync(team);
            if (tid == 0) {
                sm.cur_seq = seq;
                const long long so = J.screen_off ? J.screen_off[seq] : -1;
                sm.tbl = so >= 0 ? J.screen_bits + so : nullptr;
            }
        }
        if (wlen < min_wlen || wlen > WCAP) continue;       // long or degenerate windows: scan_generic_kernel
#ifdef RMD_STATS
        const long long ck0 = clock64();
#endif
        const long long gbase = J.seq_off[seq] + start;
        const long long b0 = J.bpp_off[w], b1 = J.bpp_off[w + 1];
        const int nnz = (int)(b1 - b0);
        const double *vals = J.bpp_p + b0;                    // values stay in global memory (L1)
        const uint32_t win_local = (uint32_t)(w - w_base);
        if (tid == 0) { sm.win_local = win_local; sm.nnz_s = nnz; }
        // the values are first needed after the staging pass: have their lines on the way (one 128-byte line per thread)
        if (!FUSED) {
UNROLL_SMALL
            for (int k = tid * 16; k < nnz; k += SCAN_THREADS * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + k));
        }

        // ---- stage the window: letters, GC table, bit matrix with ranks -----------------------------
        // (what runs once per window is kept small — no unrolling: 32 warps in different phases share 32 KB of instruction cache)
UNROLL_SMALL
        for (int k = tid; k < Smem::UW / 4; k += SCAN_THREADS) reinterpret_cast<uint4 *>(sm.Ub)[k] = make_uint4(0, 0, 0, 0);
UNROLL_SMALL
        for (int k = tid; k < CB_SLOTS * Smem::UW / 4; k += SCAN_THREADS) reinterpret_cast<uint4 *>(&sm.CB[0][0])[k] = make_uint4(0, 0, 0, 0);
UNROLL_SMALL
        for (int k = tid; k < WCAP + 16; k += SCAN_THREADS) sm.letters[k] = k < wlen ? (uint8_t)get_code(J, gbase + k) : 0;
        if (tid == 0) sm.plain = 1;
        if (tid < 2) sm.chunk_ctr[tid] = 0;
        if (new_seq && tid < RMD_N_CODES) sm.gc[tid] = J.gc_log[seq * RMD_N_CODES + tid];
        if (tid < n_models) {
            const CModel &M = c_models[tid];
            sm.gate_pass[tid] = 0;
            // odometer ranges (lib/scanner.py:161-193): p0 < n0; p1 from lo (p0 when SYMMETRIC) to max(hi, lo)
            sm.bounds[tid] = (uint32_t)rows_of(wlen, M.chain_len[0]) | ((uint32_t)max(wlen - M.chain_len[1] - 1, 0) << 8) |
                             (M.symmetric ? 1u << 16 : 0u);
        }
        team_sync(team);
        for (int k = tid; k < sm.trie_end[n_models - 1]; k += SCAN_THREADS) {
            const uint32_t bd = sm.bounds[(cinfo[k] >> 14) & 15u];    // n0 | hi << 8 | symmetric << 16, n0 >= 1
            sm.nodey[k] = ((bd & 0xffu) - 1u + 0x800u) | ((((bd >> 8) & 0xffu) + 0x800u) << 12);
        }
        if (tid < WCAP / 16 + 2) {                          // two-bit letters for the scoring engine
            uint32_t word = 0, other = 0;
UNROLL_SMALL
            for (int k = 0; k < 16; k++) {
                const int pos = tid * 16 + k;
                const uint32_t c = pos < WCAP + 16 ? sm.letters[pos] : 0;
                word |= (c & 3u) << (2 * k);
                other |= c >> 2;
            }
            sm.l2[tid] = word;
            if (other) atomicAnd(&sm.plain, 0u);
        }
        if (OTHER && tid >= 32 && tid < 32 + WCAP / 32 + 3) {   // which letters are not ACGU, one bit each
            uint32_t word = 0;
            for (int k = 0; k < 32; k++) {
                const int pos = (tid - 32) * 32 + k;
                if (pos < WCAP + 16 && sm.letters[pos] > 3) word |= 1u << k;
            }
            sm.lo[tid - 32] = word;
        }
        // bits and ranks in one pass over the list: the entry that opens a word (its predecessor lies in another
        // word) is that word's rank; a list that is not strictly increasing in (i, j), or leaves the window's upper
        // triangle, is refused
        for (int e0 = 0; e0 < nnz; e0 += SCAN_THREADS * 4) { // four entries per thread in flight
            int ei[4], ej[4], pi[4], pj[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int e = e0 + k * SCAN_THREADS + tid;
                ei[k] = ej[k] = 0; pi[k] = pj[k] = -1;
                if (FUSED) {                                  // compact transport: i | j << 8, and the value is rebuilt here
                    if (e < nnz) {
                        const uint32_t w16 = __ldcg(J.bpp_ij + b0 + e), v = __ldcg(J.bpp_q + b0 + e);   // L2: the copy engine may have just written them
                        ei[k] = w16 & 0xffu; ej[k] = w16 >> 8;
                        const double x = transport_root(v & 0x3fffffffu);                 // as expand_bpp_kernel (post_kernels.cuh)
                        const_cast<double *>(vals)[e] = __longlong_as_double(__double_as_longlong(__dmul_rn(x, x)) + (long long)(v >> 30) - 1);
                    }
                    if (e < nnz && e > 0) { const uint32_t w16 = __ldcg(J.bpp_ij + b0 + e - 1); pi[k] = w16 & 0xffu; pj[k] = w16 >> 8; }
                } else {
                    if (e < nnz) { ei[k] = J.bpp_i[b0 + e]; ej[k] = J.bpp_j[b0 + e]; }
                    if (e < nnz && e > 0) { pi[k] = J.bpp_i[b0 + e - 1]; pj[k] = J.bpp_j[b0 + e - 1]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int e = e0 + k * SCAN_THREADS + tid, i = ei[k], j = ej[k];
                if (e >= nnz) continue;
                const bool inside = i < j && j < wlen, rising = (pi[k] << 16 | pj[k]) < (i << 16 | j) || e == 0;
                if (!inside || !rising) { atomicOr(&O.raw_count[7], ERR_BAD_LIST); continue; }
                const int word = i * Smem::WORDS + (j >> 5);
                atomicOr(&sm.Ub[word], 1u << (j & 31));
                if (e == 0 || pi[k] * Smem::WORDS + (pj[k] >> 5) != word) sm.Ur[word] = (uint16_t)e;
            }
        }
        team_sync(team);
        if (!OTHER && !sm.plain) continue;                    // a letter other than ACGU: the window is the OTHER launch's (or scan_generic_kernel's)

#ifdef RMD_STATS
        const long long ck1 = clock64();
#endif
        auto lookup = [&](int a, int b) { return bpp_fast<FUSED>(sm, vals, wlen, a, b); };
        qn1 = 0; qn2 = 0; qns = 0;

        // ---- prefix screens on up to 32 gate survivors: one table bit per top-level subtree of the model's search
        // tree says whether the walk can get past the screen's last level; if none can, there is nothing to score
        auto screen_batch = [&](const int nbatch) {
            const bool act = lane < nbatch;
            const uint32_t c = act ? qs[qns - nbatch + lane] : 0;
            qns -= nbatch;
            __syncwarp();
            bool alive = act;
            const uint32_t *const tbl = sm.tbl;
            if (tbl != nullptr) {
                const uint32_t p0 = c & 0xffu, p1 = (c >> 8) & 0xffu, m = c >> 16;
                int q = sm.screen0[m];
                const int qe = sm.screen0[m + 1];
                bool screened = act && q < qe;
                if (OTHER && screened)                            // the tables are indexed by ACGU letters
                    screened = ((__funnelshift_r(sm.lo[p0 >> 5], sm.lo[(p0 >> 5) + 1], p0 & 31u) & 0xffffu) |
                                (__funnelshift_r(sm.lo[p1 >> 5], sm.lo[(p1 >> 5) + 1], p1 & 31u) << 16)) == 0;
                if (screened) {
                    const unsigned long long em = (unsigned long long)__funnelshift_r(sm.l2[p0 >> 4], sm.l2[(p0 >> 4) + 1], (p0 & 15u) * 2) |
                                                  ((unsigned long long)__funnelshift_r(sm.l2[p1 >> 4], sm.l2[(p1 >> 4) + 1], (p1 & 15u) * 2) << 32);
                    alive = false;
                    auto bit_index = [&](const int k, uint32_t &word_off) {     // letters of the screen's runs -> table bit
                        const uint4 d0 = reinterpret_cast<const uint4 *>(scr + k)[0], d1 = reinterpret_cast<const uint4 *>(scr + k)[1];
                        // d0 = word_off, shift[4], dst[4], mask[0]; d1 = mask[1..3], pad
                        uint32_t idx = ((uint32_t)(em >> (d0.y & 0xffu)) & d0.w) << (d0.z & 0xffu);
                        idx |= ((uint32_t)(em >> ((d0.y >> 8) & 0xffu)) & d1.x) << ((d0.z >> 8) & 0xffu);
                        idx |= ((uint32_t)(em >> ((d0.y >> 16) & 0xffu)) & d1.y) << ((d0.z >> 16) & 0xffu);
                        idx |= ((uint32_t)(em >> (d0.y >> 24)) & d1.z) << (d0.z >> 24);
                        word_off = d0.x + (idx >> 5);
                        return idx & 31u;
                    };
                    for (; q < qe; q += 2) {                      // two screens per step: both table reads in flight together
                        const bool two = q + 1 < qe;
                        uint32_t wa, wb;
                        const uint32_t ba = bit_index(q, wa), bb = bit_index(two ? q + 1 : q, wb);
                        const uint32_t va = __ldg(tbl + wa), vb = __ldg(tbl + wb);
                        if (((va >> ba) | (vb >> bb)) & 1u) alive = true;
                    }
                }
            }
            const uint32_t am = __ballot_sync(FULL_MASK, alive);
            if (alive) q2[qn2 + __popc(am & lt_mask)] = c;
            qn2 += __popc(am);
            __syncwarp();
        };

        // ---- gate on up to 32 queued candidates ----------------------------------------------------
        auto gate_batch = [&](const int nbatch, const int brute_model) {
            const bool act = lane < nbatch;
            const uint32_t c = act ? q1[qn1 - nbatch + lane] : 0;
            qn1 -= nbatch;
            __syncwarp();
            const int p0 = c & 0xff, p1 = (c >> 8) & 0xff, m = c >> 16;
            bool pass;
            if (BRUTE) {
                pass = gate_pass(c_models[brute_model], lookup, act, p0, p1, J.min_bpp);
            } else {
                // depth-first over the model's trie: a zero pair closes its subtree (the `break`)
                int t = act ? sm.trie0[m] : 0;
                const int tend = act ? sm.trie_end[m] : 0;
                double sum = 0.0;
                int cnt = 0;
                while (__any_sync(FULL_MASK, t < tend)) {
                    if (t < tend) {
                        const TrieNode tn = strie[t];
                        const double v = lookup(p0 + tn.ro, p1 + tn.co);
                        if (v != 0.0) { sum = __dadd_rn(sum, __dmul_rn(v, (double)tn.mult)); cnt += tn.mult; t++; }
                        else t = tn.skip;
                    }
                }
                const double mean = cnt ? __ddiv_rn(sum, (double)cnt) : 0.0;
                pass = act && mean >= J.thr_hi;
                const bool unsure = act && !pass && mean >= J.thr_lo;
                uint32_t um = __ballot_sync(FULL_MASK, unsure);
                while (um) {                                  // the reference's own sum, one model at a time
                    const int mm = __shfl_sync(FULL_MASK, m, __ffs(um) - 1);
                    const bool mine = unsure && m == mm;
                    if (gate_pass(c_models[mm], lookup, mine, p0, p1, J.min_bpp)) pass = true;
                    um &= ~__ballot_sync(FULL_MASK, mine);
                }
            }
            const uint32_t pmask = __ballot_sync(FULL_MASK, pass);
            if (pass) {
                qs[qns + __popc(pmask & lt_mask)] = c;
                atomicAdd(&sm.gate_pass[m], 1u);      // (one atomic per model through __match_any_sync measured 2 % slower)
            }
            qns += __popc(pmask);
            __syncwarp();
        };
        // Every stage below is instantiated once: the kernel's code has to stay inside the instruction cache.
        auto consume = [&](const bool drain, const int brute_model) {
            while (qn1 >= 32 || (drain && qn1 > 0)) gate_batch(min(qn1, 32), brute_model);
            if (no_score) qns = 0;
            bool again;
            do {
                again = qns >= 32 || (drain && qns > 0);
                if (again) screen_batch(min(qns, 32));
                if (qn2 >= 32 || (drain && !again)) engine(false);   // at the end of a window: until the queue is empty; walks in flight carry over
            } while (again);
        };

        if (BRUTE) {
            // every placement of the odometer is a candidate: 32 placements per lane word
            for (int m = 0; m < n_models; m++) {
                const CModel &M = c_models[m];
                const int L1 = M.chain_len[1];
                const int n0 = rows_of(wlen, M.chain_len[0]);
                const int p1_max = p1_last(wlen, L1, M.symmetric ? n0 - 1 : 0);
                const int nw = (p1_max >> 5) + 1;
                const int items = n0 * nw;
                for (int base = wid * 32; base < items; base += SCAN_THREADS) {
                    const int t = base + lane;
                    uint32_t cw = 0;
                    int p0 = 0, wd = 0;
                    if (t < items) {
                        p0 = t / nw; wd = t - p0 * nw;
                        const int lo = M.symmetric ? p0 : 0;     // odometer range of p1 for this p0
                        const int hi = p1_last(wlen, L1, lo);
                        const int blo = lo - (wd << 5), bhi = hi - (wd << 5);
                        cw = bhi < 0 || blo > 31 ? 0u : 0xffffffffu;
                        if (blo > 0 && blo <= 31) cw &= 0xffffffffu << blo;
                        if (bhi >= 0 && bhi < 31) cw &= 0xffffffffu >> (31 - bhi);
                    }
                    while (__any_sync(FULL_MASK, cw != 0)) {      // one candidate per lane per round
                        const bool has = cw != 0;
                        const int b = has ? __ffs(cw) - 1 : 0;
                        cw &= cw - 1;
                        const uint32_t pmask = __ballot_sync(FULL_MASK, has);
                        if (has) q1[qn1 + __popc(pmask & lt_mask)] = (uint32_t)p0 | ((uint32_t)((wd << 5) + b) << 8) | ((uint32_t)m << 16);
                        qn1 += __popc(pmask);
                        __syncwarp();
                        consume(false, m);
                    }
                }
                consume(true, m);                                // the exact gate reads one model's program
            }
        } else {
            // big entries x trie nodes -> placements (see the header of this file).  Every warp works on its
            // own share of the entries: no team-wide barrier in this loop.
            uint32_t *const wl = wq.wl, *const s2 = wq.s2;
            const int n_prel = c_n_prel;
            const int n_groups = c_n_groups;
            int nl = 0, ns = 0;                                  // warp-uniform fills of wl and s2
            for (int grp = 0; grp < n_groups; grp++) {
                if (grp > 0) {                                   // bitmaps of the next group of models
                    team_sync(team);
UNROLL_SMALL
                    for (int k = tid; k < CB_SLOTS * Smem::UW / 4; k += SCAN_THREADS) reinterpret_cast<uint4 *>(&sm.CB[0][0])[k] = make_uint4(0, 0, 0, 0);
                    if (tid == 0) sm.chunk_ctr[(grp + 1) & 1] = 0;
                    team_sync(team);
                }
                const int kb = c_cgrp[grp][0], ks = c_cgrp[grp][1], ke = c_cgrp[grp][2];      // the group's nodes in candidate order; [ks, ke): SYMMETRIC models
                const bool last_group = grp + 1 == n_groups;
                for (bool more = true;;) {
                    // rows shrink towards the end of the list so that the four warps finish together
                    const int want = (int)sm.chunk_ctr[grp & 1] + TAIL_ENTRIES < sm.nnz_s ? 16 : TAIL_ROW;
                    while (nl < want && more) {                   // the team's next 32 entries: keep the big ones
                        int chunk = 0;                            // taken from a team counter: the four warps stay balanced
                        if (lane == 0) chunk = (int)atomicAdd(&sm.chunk_ctr[grp & 1], 32u);
                        chunk = __shfl_sync(FULL_MASK, chunk, 0);
                        if (chunk >= sm.nnz_s) { more = false; break; }
                        const int e = chunk + lane;
                        uint32_t en = 0;
                        bool isbig = false;
                        if (e < sm.nnz_s) {
                            const double p = FUSED ? vals[e] : __ldg(vals + e);   // independent loads: one memory latency per chunk
                            uint32_t ei, ej;
                            if (FUSED) { const uint32_t w16 = __ldcg(J.bpp_ij + b0 + e); ei = w16 & 0xffu; ej = w16 >> 8; }
                            else { ei = J.bpp_i[b0 + e]; ej = J.bpp_j[b0 + e]; }
                            if (!(p >= 0.0)) atomicOr(&O.raw_count[7], ERR_BAD_VALUE);
                            if (p >= J.big_min) {
                                isbig = true;
                                // the value in units of the threshold, rounded down (at most 254): compared with `need` per node
                                const uint32_t level = (uint32_t)fmin(p * J.inv_sure_unit, 254.0);
                                en = ei | (ej << 8) | (level << 16);
                            }
                        }
                        const uint32_t bm = __ballot_sync(FULL_MASK, isbig);
                        if (isbig) wl[nl + __popc(bm & lt_mask)] = en;
                        nl += __popc(bm);
                        __syncwarp();
                    }
                    const bool flush = nl == 0;                   // the group's entries are used up: empty s2 (its bitmaps belong to this group)
                    // a row of items: lane = (entry, orientation)
                    const int take = min(nl, 16);
                    const bool valid = (lane >> 1) < take;
                    const uint32_t en = wl[valid ? nl - take + (lane >> 1) : 0];
                    nl -= take;
                    __syncwarp();
                    const int er = (lane & 1) ? (int)((en >> 8) & 0xff) : (int)(en & 0xff);
                    const int ec = (lane & 1) ? (int)(en & 0xff) : (int)((en >> 8) & 0xff);
                    uint32_t PM = valid ? 0x80000000u : 0u;       // which parent-relative pairs are present (bit 31: no parent; no item: none)
                    if (!flush && valid)
                        for (int k = 0; k < n_prel; k++) {
                            const uint32_t d = sm.prel[k];
                            if (bpp_present(sm, wlen, er + (int)(int8_t)(d & 0xff), ec + (int)(int8_t)(d >> 8))) PM |= 1u << k;
                        }
                    // the item's pair as two guarded 12-bit fields, biased by 16 so that subtracting a node's offsets never borrows
                    const uint32_t ppg = (uint32_t)(er + 16 + 0x800) | ((uint32_t)(ec + 16 + 0x800) << 12);
                    const uint32_t level = (en >> 16) & 0xffu;
                    // A node turns the item into the placement (p0, p1) = (er - ro, ec - co).  With guarded fields: ppg - x holds
                    // 0x800 + p in each field (nothing borrows across fields), b = that without the guard bits (a negative p
                    // leaves b >= 0x7f0), and c = y - b keeps both of ITS guard bits exactly when 0 <= p0 <= n0 - 1 and
                    // 0 <= p1 <= hi.  The bits above the fields (parent index, flags, need) only see borrows that go upwards.
                    int k = flush ? ke : kb;
                    for (;;) {
                        // Warp-uniform nodes, two per step (independent loads and tests): a (item, node) pair is kept when the
                        // node's parent pair is present (a root has none) and the placement lies in the odometer's range.  The
                        // pairs of a row are compacted into s2 and go on in full batches of 32: a stage that took every root
                        // node's placements straight from the row, skipping the queue, ran with 18 of 32 lanes and was 2 % slower.
                        auto push2 = [&](const bool oka, const bool okb, const uint32_t ba, const uint32_t bb, const uint4 &na, const uint4 &nb) {
                            const uint32_t oma = __ballot_sync(FULL_MASK, oka), omb = __ballot_sync(FULL_MASK, okb);
                            if (oma | omb) {
                                const int na_cnt = __popc(oma);
                                if (oka) s2[ns + __popc(oma & lt_mask)] = ba | na.z | (level >= na.w ? 0x80000000u : 0u);
                                if (okb) s2[ns + na_cnt + __popc(omb & lt_mask)] = bb | nb.z | (level >= nb.w ? 0x80000000u : 0u);
                                ns += na_cnt + __popc(omb);
                            }
                        };
                        while (k < ks && ns < 32) {               // models without SYMMETRIC: both ranges in one test
                            const bool two = k + 1 < ks;
                            const uint4 na = cnw[k], nb = cnw[two ? k + 1 : k];
                            const uint32_t ya = sm.nodey[k], yb = sm.nodey[two ? k + 1 : k];
                            const uint32_t ba = (ppg - na.x) & 0x7ff7ffu, bb = (ppg - nb.x) & 0x7ff7ffu;
                            const bool oka = (PM & na.y) && (~(ya - ba) & NODE_GUARD) == 0;
                            const bool okb = two && (PM & nb.y) && (~(yb - bb) & NODE_GUARD) == 0;
                            push2(oka, okb, ba, bb, na, nb);
                            k += two ? 2 : 1;
                        }
                        while (k >= ks && k < ke && ns < 32) {    // SYMMETRIC (lib/scanner.py:183-188): p1 runs from p0 to max(hi, p0)
                            const uint4 na = cnw[k];
                            const uint32_t ya = sm.nodey[k];
                            const uint32_t ba = (ppg - na.x) & 0x7ff7ffu;
                            const uint32_t q0 = ba & 0xfffu, q1 = ba >> 12, hi = (ya >> 12) & 0x7ffu;
                            const bool oka = (PM & na.y) && ((ya - ba) & 0x800u) && q1 >= q0 && (q1 <= hi || q1 == q0);
                            push2(oka, false, ba, ba, na, na);
                            k++;
                        }
                        const bool row_done = flush && k >= ke;   // the closing pass of the group: whatever is left in s2
                        if (ns < 32 && !row_done) break;          // every node has seen this row
                        const int n = min(ns, 32);
                        bool ok = lane < n;
                        __syncwarp();                             // the pairs pushed above are visible to the lanes that take them
                        const uint32_t it = ok ? s2[ns - n + lane] : 0;
                        ns -= n;
                        __syncwarp();
                        const uint32_t plc = it & 0xffffffu;      // placement p0 | p1 << 12
                        bool sure = it >> 31;
                        const uint32_t info = cinfo[(it >> 24) & 0x7fu];   // CandNode::info of the node that made it
                        // every pair before the parent must be non-zero too (the reference's `break`)
                        uint32_t a = (info >> 7) & 0x7fu;
                        while (__any_sync(FULL_MASK, ok && a != 0x7fu)) {
                            if (ok && a != 0x7fu) {
                                const uint32_t w = ancw[a];
                                ok = bpp_present_f(sm, plc + (w & 0xffffffu));
                                a = w >> 24;
                            }
                        }
                        const uint32_t p0 = plc & 0xfffu, p1 = plc >> 12;
                        if (ok) {                                 // a placement is queued once
                            const uint32_t slot1 = (info >> 18) & 3u;
                            if (slot1) {
                                const uint32_t bit = 1u << (p1 & 31);
                                ok = !(atomicOr(&sm.CB[slot1 - 1][p0 * Smem::WORDS + (p1 >> 5)], bit) & bit);
                            } else if (info & (1u << 20)) {
                                // a trie without a bitmap (two nodes): its first node produces the placement whenever that node's
                                // own pair is big, the other node only otherwise (one entry and one node give one placement)
                                const uint32_t f = plc + (ancw[(info >> 21) & 0x7fu] & 0xffffffu);
                                ok = !(lookup((int)(f & 0xfffu), (int)(f >> 12)) >= J.big_min);
                            }
                        }
                        sure = sure && ok;
                        const uint32_t om = __ballot_sync(FULL_MASK, ok);
                        const uint32_t smask = __ballot_sync(FULL_MASK, sure), cmask = om & ~smask;
                        const uint32_t model = (info >> 14) & 15u;
                        const uint32_t item = p0 | (p1 << 8) | (model << 16);
                        if (sure) {
                            qs[qns + __popc(smask & lt_mask)] = item;
                            atomicAdd(&sm.gate_pass[model], 1u);
                        } else if (ok) q1[qn1 + __popc(cmask & lt_mask)] = item;
                        qns += __popc(smask);
                        qn1 += __popc(cmask);
                        __syncwarp();
                        consume(row_done && last_group, 0);
                        if (row_done) break;
                    }
                    if (flush) break;
                }
            }
        }
#ifdef RMD_STATS
        const long long ck2 = clock64();
#endif
        if (OTHER) engine(true);                               // the walks read this window's letters: finish them here
#ifdef RMD_STATS
        const long long ck3 = clock64();
#endif
        team_sync(team);
        if (tid < n_models) {
            O.gate_pass[(long long)win_local * n_models + tid] = sm.gate_pass[tid];
            if (sm.gate_pass[tid]) atomicAdd(&O.raw_count[2], (unsigned long long)sm.gate_pass[tid]);
        }
        if (tid == 0) atomicAdd(&O.raw_count[1], odometer_count_all(n_models, wlen));   // tests_all: closed form of the odometer
#ifdef RMD_STATS
        if (lane == 0) {
            atomicAdd(&O.raw_count[4], dbg_visits); atomicAdd(&O.raw_count[5], dbg_iter); dbg_visits = 0; dbg_iter = 0;
            const long long ck4 = clock64();
            atomicAdd(&O.stats[0], (unsigned long long)(ck1 - ck0)); atomicAdd(&O.stats[1], (unsigned long long)(ck2 - ck1));
            atomicAdd(&O.stats[2], (unsigned long long)(ck3 - ck2)); atomicAdd(&O.stats[3], (unsigned long long)(ck4 - ck3));
        }
#endif
    }
}

// Windows of [w_begin, w_end) that the shared-memory kernel takes (length within [min_wlen, max_wlen]) and that hold a
// letter other than ACGU: the OTHER launch's work list (window numbers relative to w_begin, any order).  One warp per window.
__global__ void __launch_bounds__(256) other_windows_kernel(const DevJob J, const long long w_begin, const long long w_end, const int min_wlen,
                                                            const int max_wlen, uint32_t *list, uint32_t *count) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = w_begin + warp; w < w_end; w += n_warps) {
        int seq, wlen;
        long long start;
        window_geom(J, w, seq, start, wlen);
        if (wlen < min_wlen || wlen > max_wlen) continue;
        const uint8_t *c = J.codes8 + J.seq_off[seq] + start;
        bool other = false;
        for (int k = lane; k < wlen; k += 32) other |= c[k] > 3;
        if (__any_sync(FULL_MASK, other) && lane == 0) list[atomicAdd(count, 1u)] = (uint32_t)(w - w_begin);
    }
}

// ------------------------------------------------------------------------------------------------
// Generic path for windows longer than FAST_W (large --win-len, or whole-sequence mode,
// lib/scanner.py:43-44).  One CTA per (window, model); placements are visited in odometer order in
// chunks of 256, the bpp list is probed by binary search in global memory.  Exact, not fast.
// ------------------------------------------------------------------------------------------------
#define GEN_THREADS 256
#define GEN_BITS_W 512

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
                    const ScanOut O, const int min_fast_wlen, const int max_fast_wlen, const int fast_takes_other) {
    __shared__ uint32_t scan_tmp[GEN_THREADS / 32 + 1];
    __shared__ double gc[RMD_N_CODES];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_pass;
    // which pairs are stored, for windows up to GEN_BITS_W letters: bit (lo * GEN_BITS_W + hi).  Nearly every pair the gate asks
    // for is absent; the bit answers that without the binary search.
    __shared__ uint32_t s_bits[GEN_BITS_W * GEN_BITS_W / 32];
    const int tid = threadIdx.x;
    const long long grp_local = blockIdx.x;
    const int m = (int)(grp_local % J.n_models);
    const long long w = w_begin + grp_local / J.n_models;
    if (w >= w_end) return;
    int seq, wlen;
    long long start;
    window_geom(J, w, seq, start, wlen);
    if (wlen <= 0) return;
    if (wlen >= min_fast_wlen && wlen <= max_fast_wlen) {                         // scan_fast_kernel's, unless a letter is not ACGU and no OTHER launch follows
        if (fast_takes_other) return;
        int other = 0;
        if (J.codes8) for (int k = tid; k < wlen; k += GEN_THREADS) other |= J.codes8[J.seq_off[seq] + start + k] > 3;
        if (!__syncthreads_or(other)) return;
    }
    const CModel &M = c_models[m];
    const long long gbase = J.seq_off[seq] + start;
    const long long b0 = J.bpp_off[w];
    const int nnz = (int)(J.bpp_off[w + 1] - b0);
    if (nnz == 0 && J.min_bpp > 0.0) {             // no pair, positive threshold: every placement fails the gate (mean 0.0)
        if (tid == 0) {
            const long long grp = (w - w_base) * J.n_models + m;
            O.gate_pass[grp] = 0;
            O.group_hits[grp] = 0;
            atomicAdd(&O.raw_count[1], odometer_count(M, wlen));
        }
        return;
    }
    if (tid < RMD_N_CODES) gc[tid] = J.gc_log[seq * RMD_N_CODES + tid];
    if (tid == 0) s_pass = 0;
    const bool bits = wlen <= GEN_BITS_W;
    if (bits) {
        for (int k = tid; k < GEN_BITS_W * GEN_BITS_W / 32; k += GEN_THREADS) s_bits[k] = 0;
        __syncthreads();
        for (int e = tid; e < nnz; e += GEN_THREADS) {
            const int i = J.bpp_i[b0 + e], j = J.bpp_j[b0 + e];
            if (i < j && j < wlen) atomicOr(&s_bits[(i * GEN_BITS_W + j) >> 5], 1u << (j & 31));
        }
    }
    __syncthreads();
    auto lookup = [&](int a, int b) {
        if (bits) {
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            if (lo < 0 || hi >= wlen || !((s_bits[(lo * GEN_BITS_W + hi) >> 5] >> (hi & 31)) & 1u)) return 0.0;
        }
        return bpp_search(J, b0, nnz, wlen, a, b);
    };
    auto letter = [&](int pos) { return get_code(J, gbase + pos); };
    const int L0 = M.chain_len[0], L1 = M.chain_len[1];
    const int n0 = rows_of(wlen, L0);
    unsigned long long tests_all = 0;
    uint32_t group_rank = 0, npass = 0;
    // one placement per thread: gate, walk, records appended through a block-wide scan
    auto try_placements = [&](const bool act, const int p0, const int p1) {
        const bool pass = gate_pass(M, lookup, act, p0, p1, J.min_bpp);
        EvalOut r; r.leaf = -1;
        // a window shorter than a chain still has its placement (0, 0) tried (lib/scanner.py:137-141); scoring it would
        // read letters past the window's end, where the reference raises IndexError: such a placement yields no hit
        if (pass) { npass++; if (p0 + L0 <= wlen && p1 + L1 <= wlen) r = eval_placement(M, letter, gc, p0, p1, J.cutoff); }
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
                h.rank = group_rank + rank;                          // (the ordering pass ranks records by (p0, p1) itself)
                h.pm = r.pm; h.pr = r.pr;
                O.raw[slot] = h;
            }
            group_rank += total;
            __syncthreads();
        }
    };
    if (J.min_bpp > 0.0 && !M.brute) {
        // Placements from the list, not from the odometer: with a positive threshold the mean is positive only if some
        // configuration sums at least its first pair (lib/scanner.py:204-206), i.e. a root pair of the gate trie is stored
        // and non-zero.  Every stored entry, in both orientations, with every root gives one placement; when several roots
        // have their pair at a placement the first of them owns it.  2 * nnz * roots gate evaluations instead of one per
        // placement of the odometer (a 300-letter window: 20 000 instead of 336 000 for the four shipped models).
        tests_all = odometer_count(M, wlen);
        const int t0 = M.trie0, t1 = M.trie0 + M.n_trie;
        for (long long item0 = 0; item0 < 2LL * nnz; item0 += GEN_THREADS) {
            const long long item = item0 + tid;
            int er = 0, ec = 0;
            bool have = false;
            if (item < 2LL * nnz) {
                const long long e = b0 + (item >> 1);
                const int i = J.bpp_i[e], j = J.bpp_j[e];
                have = J.bpp_p[e] != 0.0;                            // a stored zero counts as absent (lib/scanner.py:205)
                er = (item & 1) ? j : i; ec = (item & 1) ? i : j;
            }
            for (int r = t0; r < t1; r++) {                          // block-uniform
                const TrieNode tn = c_trie[r];
                if (tn.parent >= 0) continue;
                const int p0 = er - tn.ro, p1 = ec - tn.co;
                const int lo = M.symmetric ? p0 : 0;
                bool act = have && p0 >= 0 && p0 < n0 && p1 >= lo && p1 <= p1_last(wlen, L1, lo);
                if (act)
                    for (int q = t0; q < r; q++)
                        if (c_trie[q].parent < 0 && lookup(p0 + c_trie[q].ro, p1 + c_trie[q].co) != 0.0) { act = false; break; }
                try_placements(act, p0, p1);
            }
        }
    } else {
        for (int p0 = 0; p0 < n0; p0++) {
            const int lo = M.symmetric ? p0 : 0, hi = p1_last(wlen, L1, lo);
            tests_all += (unsigned long long)(hi - lo + 1);
            for (int c0 = lo; c0 <= hi; c0 += GEN_THREADS) try_placements(c0 + tid <= hi, p0, c0 + tid);
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
