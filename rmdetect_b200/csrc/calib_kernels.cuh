// calib_kernels.cuh — K5: rmcalibrate's background scoring.
//
// Restates rmcalibrate.py:265-295 (reference root) for a batch of samples:
//   :276      sequence = s[:L0] + "CCCCCC" + s[L0:], pointers = [0, L0 + 6]   (:229)
//   :279-281  model.root.eval(...) with GC.neutral(); get_solution()
//   :47-62    get_gc_scores: score_i = max((prob_model - sum_{used letters} log gc_i[c]) / log 2, 0)
//             over the letters the winning gap configuration consumed, chain 0 then chain 1
//   :35-45    get_gc_weights: weight_i = e ** (sum_{all L letters} log gc_i[c] + L log 4)
//   :292-295  hist_i[round(score_i, 1)] += weight_i
// Buckets are the reference's exactly (scores near a bucket boundary are recomputed in its order, see below); histogram
// sums are order-free fp64 sums (relative difference ~1e-13, far inside the 1e-6 bound north_star states).
//
// Three kernels.  A sample that reaches no leaf of the search tree (all but a few in a thousand) scores 0 in every class
// and only adds its weight to bucket 0, and the weight depends on the sample only through its letter counts
// (n_A, n_C, n_G, n_U), which add up to L:  weight_i = (4 g_iU)^L x prod_{k in ACG} (g_ik / g_iU)^(n_k).  So
//   calib_table_kernel   tabulates the powers once per session (n_class x 3 x (L + 1) doubles, the constant factor
//                        folded into the first of the three);
//   calibrate_kernel     one thread per sample: draw (or read) it, walk the tree, and count it in the cell
//                        (n_A, n_C, n_G) of a per-CTA histogram in shared memory — one atomic instead of 165 weights;
//                        a sample with a leaf is set aside in a list;
//   calib_ok_kernel      the listed samples, one thread per (sample, class): the reference's per-class loop (score,
//                        bucket, weight) with full warps (inside calibrate_kernel it took most of the time: a warp
//                        with one such lane ran 165 iterations for it);
//   calib_cells_kernel   bucket 0 of class i += sum over the cells of count x weight_i(cell): 165 x 1771 products
//                        instead of 165 per sample.
// A table entry is an `exp` of an exponent with one rounding, so a weight is within 1e-14 of the reference's (the tests
// ask for 1e-9).  (Round 1/2a: every thread computed 165 weights with an `exp` each and the warps reduced bucket 0 by
// shuffles: 0.78 ms per 2^20 KT samples.)
#pragma once
#include "device_common.cuh"

#ifndef CAL_THREADS
#define CAL_THREADS 512                 // 2 x 512 threads per SM at 64 registers (measured: 4 x 256 is 5 % slower, 3 x 256 at 85 registers 7 %)
#define CAL_CTAS_PER_SM 2
#endif
#define CAL_MAX_CLASS 256
#define CAL_SPACER 6
#ifndef CAL_PIECE
#define CAL_PIECE (1LL << 22)           // samples per launch: rmd_calib_add cuts a larger batch into pieces
#endif
#define CAL_SMEM_L 20                   // samples up to 20 letters (every shipped model) count in shared memory: 21^3 cells, 37 KB

struct CalibOk { unsigned long long bits, used; double pm; int n_used, pad; };   // sample, consumed letters (2 bits each), prob_model

struct CalibArgs {
    int model, L, n_class, n_bins;
    long long n, first;
    const uint8_t *samples;     // [n][L] codes, or nullptr: draw on the device
    unsigned long long seed;
    const double *class_log;    // [n_class][4]
    const double *wtab;         // [n_class][3][L + 1] powers of the class frequencies (calib_table_kernel)
    uint32_t *cells;            // [(L + 1)^3] samples without a solution, by letter counts; then the number of samples with one (zeroed per batch)
    struct CalibOk *ok;         // the samples with a solution of the current batch
    const XNode *xnodes;        // the scoring engine's tables (all models; CModel::xnode0)
    const double *xcpt;
    const double *thr;          // [n_bins] rounding thresholds
    double cutoff, ln2, K;      // K = L * log(4.0)
    double *hist;               // [n_class][n_bins]
    unsigned long long *n_ok;
    int *overflow;
};

__device__ __forceinline__ unsigned long long mix64_dev(unsigned long long x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// The table of powers for one session: tab[i][k][n], k = 0..2 (A, C, G against U), n = 0..L
__global__ void __launch_bounds__(256) calib_table_kernel(const double *class_log, const int n_class, const int L, double *tab) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x, NP = L + 1;
    if (idx >= n_class * 3 * NP) return;
    const int i = idx / (3 * NP), k = (idx / NP) % 3, n = idx % NP;
    const double lu = class_log[4 * i + 3];
    double e = (double)n * (class_log[4 * i + k] - lu);
    if (k == 0) e += (double)L * (lu + 1.3862943611198906);          // (4 g_U)^L
    tab[idx] = exp(e);
}

// the reference's own arithmetic for a score too close to a bucket boundary to call from the letter counts
__device__ __noinline__ int calib_exact_bin(const double *lg, const unsigned long long used, const int n_used, const double pm,
                                            const double ln2, const double *thr, const int n_bins) {
    double p = 0.0;
    for (int k = 0; k < n_used; k++) p = __dadd_rn(p, lg[(used >> (2 * k)) & 3]);
    double score = __ddiv_rn(__dsub_rn(pm, p), ln2);
    if (!(score > 0.0)) return 0;
    int bin = (int)(score * 10.0 + 0.5);
    if (bin > n_bins) bin = n_bins;
    while (bin > 0 && score < __ldg(thr + bin - 1)) bin--;
    while (bin < n_bins && score >= __ldg(thr + bin)) bin++;
    return bin;
}

// a sample with a solution goes to the list of calib_ok_kernel with the letters its winning configuration consumed
__device__ __noinline__ void calib_set_aside(const CalibArgs &A, const CModel &M, const unsigned long long bits, const int leaf,
                                             const double pm, const int L, const int L0, const int n_cells) {
    CalibOk rec;
    rec.bits = bits; rec.used = 0; rec.pm = pm; rec.n_used = 0; rec.pad = 0;
    const int8_t *lo = M.leaf_offsets + leaf * L;
    for (int k = 0; k < L; k++) {                     // chain 0 then chain 1
        const int o = lo[k];
        if (o >= 0) {
            const int pos = k < L0 ? o : L0 + o;      // index into the sample
            rec.used |= ((bits >> (2 * pos)) & 3ULL) << (2 * rec.n_used);
            rec.n_used++;
        }
    }
    A.ok[atomicAdd(A.cells + n_cells, 1u)] = rec;
}

// One thread per sample; persistent CTAs (grid-stride over batches of CAL_THREADS samples) with the count histogram in
// shared memory (SMEM_CELLS: (L + 1)^3 cells fit) or straight in global memory (samples longer than CAL_SMEM_L letters).
template <bool SMEM_CELLS>
__global__ void __launch_bounds__(CAL_THREADS, CAL_CTAS_PER_SM) calibrate_kernel(const CalibArgs A) {
    extern __shared__ uint32_t cal_cells[];           // [(L + 1)^3], index (n_A * (L + 1) + n_C) * (L + 1) + n_G
    const int tid = threadIdx.x;
    const int L = A.L, NP = L + 1, n_cells = NP * NP * NP;
    if (SMEM_CELLS)
        for (int k = tid; k < n_cells; k += CAL_THREADS) cal_cells[k] = 0;
    __syncthreads();

    const CModel &M = c_models[A.model];
    const int L0 = M.chain_len[0];
    double neutral[RMD_N_CODES];
#pragma unroll
    for (int k = 0; k < RMD_N_CODES; k++) neutral[k] = -1.3862943611198906;   // math.log(0.25), GC.neutral()
    for (long long s = (long long)blockIdx.x * CAL_THREADS + tid; s < A.n; s += (long long)gridDim.x * CAL_THREADS) {
        unsigned long long bits = 0;                  // the sample, 2 bits per letter
        if (A.samples) {
            for (int k = 0; k < L; k++) bits |= (unsigned long long)(A.samples[s * L + k] & 3) << (2 * k);
        } else {
            // sample s of the stream: the low 2L bits of one 64-bit hash (host twin: synth.synth_calibration_samples)
            const unsigned long long h = mix64_dev(A.seed + 0x9E3779B97F4A7C15ULL * ((unsigned long long)(A.first + s) + 1ULL));
            bits = L >= 32 ? h : h & ((1ULL << (2 * L)) - 1ULL);
        }
        auto letter = [&](int pos) {                  // the sample with the CCCCCC spacer spliced in
            if (pos < L0) return (int)((bits >> (2 * pos)) & 3);
            if (pos < L0 + CAL_SPACER) return 1;
            return (int)((bits >> (2 * (pos - CAL_SPACER))) & 3);
        };
        const EvalOut r = eval_placement(M, letter, neutral, 0, L0 + CAL_SPACER, A.cutoff);
        // letter counts of the whole sample: A = 0, C = 1, G = 2, U = 3 in two bits per letter
        const unsigned long long lo1 = bits & 0x5555555555555555ULL, hi1 = (bits >> 1) & 0x5555555555555555ULL;
        const int n_c = __popcll(lo1 & ~hi1), n_g = __popcll(hi1 & ~lo1), n_a = L - n_c - n_g - __popcll(lo1 & hi1);
        if (r.leaf < 0) {                             // no solution: score 0 in every class, the weights follow from the counts
            const int cell = (n_a * NP + n_c) * NP + n_g;
            if (SMEM_CELLS) atomicAdd(&cal_cells[cell], 1u);
            else atomicAdd(A.cells + cell, 1u);
            continue;
        }
        // a sample with a solution (a few in a thousand) scores in every class: it is set aside for calib_ok_kernel,
        // which works through (sample, class) pairs with full warps
        calib_set_aside(A, M, bits, r.leaf, r.pm, L, L0, n_cells);
    }
    __syncthreads();
    if (SMEM_CELLS)
        for (int k = tid; k < n_cells; k += CAL_THREADS)
            if (cal_cells[k]) atomicAdd(A.cells + k, cal_cells[k]);
}

// The same with the scan kernel's scoring engine (scan_kernels.cuh; lib/node.py:176-277 on the pre-decoded 32-byte nodes):
// walks of random samples end after 2 to 272 tree nodes, and with one sample per thread a warp waited for its longest walk
// (7 of 32 lanes active, ncu).  Here a warp owns a contiguous range of the batch, every lane carries its own sample and walk,
// visits one tree node per iteration, and lanes that have finished take the warp's next samples once CAL_REFILL_IDLE of
// them are idle.  For model sets the engine's node format holds (chains up to 16 letters) and a cutoff that lets the root pass.
#ifndef CAL_REFILL_IDLE
#define CAL_REFILL_IDLE 8
#endif
template <bool SMEM_CELLS>
__global__ void __launch_bounds__(CAL_THREADS, CAL_CTAS_PER_SM) calibrate_engine_kernel(const CalibArgs A) {
    extern __shared__ uint32_t cal_cells[];           // [(L + 1)^3], index (n_A * (L + 1) + n_C) * (L + 1) + n_G
    const int tid = threadIdx.x, lane = tid & 31;
    const int L = A.L, NP = L + 1, n_cells = NP * NP * NP;
    if (SMEM_CELLS)
        for (int k = tid; k < n_cells; k += CAL_THREADS) cal_cells[k] = 0;
    __syncthreads();
    uint32_t lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const CModel &M = c_models[A.model];
    const int L0 = M.chain_len[0];
    const uint32_t xnode0 = (uint32_t)M.xnode0;
    const uint4 *const xn = reinterpret_cast<const uint4 *>(A.xnodes);
    const double g = -1.3862943611198906;             // math.log(0.25): GC.neutral() gives every letter (and the gap) this
    const int d = L0 + CAL_SPACER;                    // pointers [0, L0 + 6] (rmcalibrate.py:229)
    const unsigned long long m0 = (1ULL << (2 * L0)) - 1ULL;   // chain 0 of the sample (chains hold at most 16 letters)
    // this warp's samples
    const long long n_warps = (long long)gridDim.x * (CAL_THREADS / 32), warp = (long long)blockIdx.x * (CAL_THREADS / 32) + (tid >> 5);
    const long long per = (A.n + n_warps - 1) / n_warps;
    long long next = min(A.n, warp * per);
    const long long end = min(A.n, next + per);

    uint32_t e_t = XNODE_END;                         // node of the lane's walk (XNODE_END: idle)
    unsigned long long bits = 0, e_m = 0;             // the sample; its chains as the engine reads them (chain 1 from bit 32 on)
    double e_pm = 0.0, e_pr = 0.0, e_bpm = -500.0;
    int e_leaf = -1;
    double2 e_stk[RMD_MAX_BRANCH_DEPTH + 2];          // sums saved at gap branches (slot 0: the implicit root's zeros)
    e_stk[0] = make_double2(0.0, 0.0);
    for (;;) {
        const uint32_t im = __ballot_sync(FULL_MASK, e_t == XNODE_END);
        if (next < end && (__popc(im) >= CAL_REFILL_IDLE || im == FULL_MASK)) {   // idle lanes take the next samples
            const long long s = next + __popc(im & lt_mask);
            if ((im >> lane & 1) && s < end) {
                if (A.samples) {
                    bits = 0;
                    for (int k = 0; k < L; k++) bits |= (unsigned long long)(A.samples[s * L + k] & 3) << (2 * k);
                } else {
                    const unsigned long long h = mix64_dev(A.seed + 0x9E3779B97F4A7C15ULL * ((unsigned long long)(A.first + s) + 1ULL));
                    bits = L >= 32 ? h : h & ((1ULL << (2 * L)) - 1ULL);
                }
                e_m = (bits & m0) | ((bits >> (2 * L0)) << 32);
                e_t = xnode0; e_pm = 0.0; e_pr = 0.0; e_bpm = -500.0; e_leaf = -1;     // lib/node.py:250-251
            }
            next = min(end, next + __popc(im));
        }
        const bool busy = e_t != XNODE_END;
        if (!__any_sync(FULL_MASK, busy)) break;      // nothing in flight and nothing left to take
        if (busy) {
            const uint4 a = __ldg(xn + 2 * e_t);                      // sel, cpt, m0 | m1 << 16, skip | next << 16
            const uint2 b = __ldg(reinterpret_cast<const uint2 *>(xn + 2 * e_t + 1));   // leaf | model << 16 | flags << 24 | bdepth << 28, dlo | dspan << 16
            const int nt = (a.x & XSEL_OWN_GAP) ? 4 : (int)((uint32_t)(e_m >> (a.x & 63u)) & 3u);          // :199-213
            const uint32_t c0 = (uint32_t)(e_m >> ((a.x >> 8) & 63u)) & 3u, c1 = (uint32_t)(e_m >> ((a.x >> 16) & 63u)) & 3u;
            double v = __ldg(A.xcpt + a.y + nt + c0 * (a.z & 0xffffu) + c1 * (a.z >> 16));
            if (v == CUDART_INF) v = g;                               // GC-sentinel row (:217-218)
            const uint32_t fl = b.x;
            const int bdepth = fl >> 28;
            if (fl & (RMD_F_RESTORE << 24)) { const double2 sv = e_stk[bdepth]; e_pm = sv.x; e_pr = sv.y; }
            const bool apart = (uint32_t)(d - (int)(short)(b.y & 0xffffu)) >= (b.y >> 16);   // :184-192 as one interval test
            e_pm = __dadd_rn(e_pm, v);                                // :227-228
            e_pr = __dadd_rn(e_pr, g);
            const bool alive = apart && e_pm >= __dadd_rn(e_pr, A.cutoff);             // :248
            if (alive && (fl & (RMD_F_LEAF << 24)) && e_pm > e_bpm) { e_bpm = e_pm; e_leaf = (int)(fl & 0xffffu); }   // :265-275
            if (alive && (fl & (RMD_F_SAVE << 24))) e_stk[bdepth + 1] = make_double2(e_pm, e_pr);
            e_t = alive ? (a.w >> 16) : (a.w & 0xffffu);
            if (e_t == XNODE_END) {                                   // the walk is over
                if (e_leaf < 0) {                     // no solution: score 0 in every class, the weights follow from the counts
                    const unsigned long long lo1 = bits & 0x5555555555555555ULL, hi1 = (bits >> 1) & 0x5555555555555555ULL;
                    const int n_c = __popcll(lo1 & ~hi1), n_g = __popcll(hi1 & ~lo1), n_a = L - n_c - n_g - __popcll(lo1 & hi1);
                    const int cell = (n_a * NP + n_c) * NP + n_g;
                    if (SMEM_CELLS) atomicAdd(&cal_cells[cell], 1u);
                    else atomicAdd(A.cells + cell, 1u);
                } else {
                    calib_set_aside(A, M, bits, e_leaf, e_bpm, L, L0, n_cells);
                }
            }
        }
    }
    __syncthreads();
    if (SMEM_CELLS)
        for (int k = tid; k < n_cells; k += CAL_THREADS)
            if (cal_cells[k]) atomicAdd(A.cells + k, cal_cells[k]);
}

// The samples with a solution, one thread per (sample, class): the reference's per-class loop (rmcalibrate.py:47-62,292-295)
__global__ void __launch_bounds__(256) calib_ok_kernel(const CalibArgs A) {
    const int L = A.L, NP = L + 1;
    const uint32_t cnt = A.cells[NP * NP * NP];
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt) atomicAdd(A.n_ok, (unsigned long long)cnt);
    const long long total = (long long)cnt * A.n_class;
    const double inv_ln2 = 1.0 / A.ln2;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int j = (int)(idx / A.n_class), i = (int)(idx - (long long)j * A.n_class);
        const CalibOk rec = A.ok[j];
        // letter counts of the whole sample (weight) and of the consumed letters (score): two bits per letter
        unsigned long long lo1 = rec.bits & 0x5555555555555555ULL, hi1 = (rec.bits >> 1) & 0x5555555555555555ULL;
        const int n_c = __popcll(lo1 & ~hi1), n_g = __popcll(hi1 & ~lo1), n_a = L - n_c - n_g - __popcll(lo1 & hi1);
        lo1 = rec.used & 0x5555555555555555ULL; hi1 = (rec.used >> 1) & 0x5555555555555555ULL;
        const int u_c = __popcll(lo1 & ~hi1), u_g = __popcll(hi1 & ~lo1), u_u = __popcll(lo1 & hi1), u_a = rec.n_used - u_c - u_g - u_u;
        const double *lg = A.class_log + 4 * i, *tb = A.wtab + (long long)i * 3 * NP;
        const double weight = __ldg(tb + n_a) * __ldg(tb + NP + n_c) * __ldg(tb + 2 * NP + n_g);
        // The score decides a bucket, which must be the reference's: it is first estimated from the counts (within 1e-12 of
        // the reference's letter-by-letter sum) and recomputed in the reference's order only when the estimate lies within
        // 1e-8 of a bucket boundary.
        const double est = (rec.pm - ((double)u_a * __ldg(lg) + (double)u_c * __ldg(lg + 1) + (double)u_g * __ldg(lg + 2) + (double)u_u * __ldg(lg + 3))) * inv_ln2;
        const double scaled = est * 10.0 + 0.5;
        int bin = scaled > 0.0 ? (int)scaled : 0;
        const double frac = scaled - (double)bin;
        if (scaled > -1e-6 && (frac < 1e-7 || frac > 1.0 - 1e-7))               // too close to call
            bin = calib_exact_bin(lg, rec.used, rec.n_used, rec.pm, A.ln2, A.thr, A.n_bins);
        if (bin >= A.n_bins) atomicOr(A.overflow, 1);
        else atomicAdd(A.hist + (long long)i * A.n_bins + bin, weight);
    }
}

// bucket 0 of class blockIdx.x += sum over the cells of count x weight(cell)
__global__ void __launch_bounds__(256) calib_cells_kernel(const CalibArgs A) {
    __shared__ double part[8];
    const int i = blockIdx.x, NP = A.L + 1;
    const double *tb = A.wtab + (long long)i * 3 * NP;
    double sum = 0.0;
    const int n_cells = NP * NP * NP;
    for (int k0 = threadIdx.x; k0 < n_cells; k0 += 4 * 256) {       // four counts in flight per thread (most cells are empty)
        uint32_t cnt[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cnt[u] = k0 + u * 256 < n_cells ? __ldg(A.cells + k0 + u * 256) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (cnt[u]) {
                const int k = k0 + u * 256, g = k % NP, c = (k / NP) % NP, a = k / (NP * NP);
                sum += (double)cnt[u] * (__ldg(tb + a) * __ldg(tb + NP + c) * __ldg(tb + 2 * NP + g));
            }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_down_sync(FULL_MASK, sum, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += part[w];
        if (t != 0.0) A.hist[(long long)i * A.n_bins] += t;       // one CTA per class, after calibrate_kernel on the same stream
    }
}

// first and last bucket occupied in any class: out[0] = min, out[1] = max (preset to n_bins and -1); one CTA per class
__global__ void __launch_bounds__(256) calib_cols_kernel(const double *hist, const int n_class, const int n_bins, int *out) {
    const double *row = hist + (long long)blockIdx.x * n_bins;
    int lo = n_bins, hi = -1;
    for (int b = threadIdx.x; b < n_bins; b += 256)
        if (row[b] != 0.0) { lo = min(lo, b); hi = max(hi, b); }
    lo = __reduce_min_sync(FULL_MASK, lo);
    hi = __reduce_max_sync(FULL_MASK, hi);
    if ((threadIdx.x & 31) == 0 && hi >= 0) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}
