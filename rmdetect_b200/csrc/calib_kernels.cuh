// calib_kernels.cuh — K5: rmcalibrate's background scoring.
//
// Restates rmcalibrate.py:265-295 (reference root) for a batch of samples:
//   :276      sequence = s[:L0] + "CCCCCC" + s[L0:], pointers = [0, L0 + 6]   (:229)
//   :279-281  model.root.eval(...) with GC.neutral(); get_solution()
//   :47-62    get_gc_scores: score_i = max((prob_model - sum_{used letters} log gc_i[c]) / log 2, 0)
//             over the letters the winning gap configuration consumed, chain 0 then chain 1
//   :35-45    get_gc_weights: weight_i = e ** (sum_{all L letters} log gc_i[c] + L log 4)
//   :292-295  hist_i[round(score_i, 1)] += weight_i
// One thread per sample, classes in an inner loop with the per-class logs in shared memory.
// Buckets are the reference's exactly (scores near a bucket boundary are recomputed in its order, see below);
// histogram sums are order-free fp64 atomics (relative difference ~1e-13, far inside the 1e-6 bound).
#pragma once
#include "device_common.cuh"

#define CAL_THREADS 256
#define CAL_MAX_CLASS 256
#define CAL_SPACER 6

struct CalibArgs {
    int model, L, n_class, n_bins;
    long long n, first;
    const uint8_t *samples;     // [n][L] codes, or nullptr: draw on the device
    unsigned long long seed;
    const double *class_log;    // [n_class][4]
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

__global__ void __launch_bounds__(CAL_THREADS) calibrate_kernel(const CalibArgs A) {
    __shared__ double cls[CAL_MAX_CLASS * 4];
    __shared__ double bin0[CAL_THREADS / 32][CAL_MAX_CLASS];
    __shared__ unsigned int s_ok;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = tid; k < A.n_class * 4; k += CAL_THREADS) cls[k] = A.class_log[k];
    for (int k = tid; k < (CAL_THREADS / 32) * CAL_MAX_CLASS; k += CAL_THREADS) (&bin0[0][0])[k] = 0.0;
    if (tid == 0) s_ok = 0;
    __syncthreads();

    const CModel &M = c_models[A.model];
    const int L = A.L, L0 = M.chain_len[0];
    const long long s = (long long)blockIdx.x * CAL_THREADS + tid;
    const bool live = s < A.n;
    unsigned long long bits = 0;                      // the sample, 2 bits per letter
    if (live) {
        if (A.samples) {
            for (int k = 0; k < L; k++) bits |= (unsigned long long)(A.samples[s * L + k] & 3) << (2 * k);
        } else {
            const unsigned long long g0 = (unsigned long long)(A.first + s) * (unsigned long long)L;
            for (int k = 0; k < L; k++)
                bits |= (mix64_dev(A.seed + 0x9E3779B97F4A7C15ULL * (g0 + k + 1)) >> 62) << (2 * k);
        }
    }
    auto letter = [&](int pos) {                      // the sample with the CCCCCC spacer spliced in
        if (pos < L0) return (int)((bits >> (2 * pos)) & 3);
        if (pos < L0 + CAL_SPACER) return 1;
        return (int)((bits >> (2 * (pos - CAL_SPACER))) & 3);
    };
    double neutral[RMD_N_CODES];
#pragma unroll
    for (int k = 0; k < RMD_N_CODES; k++) neutral[k] = -1.3862943611198906;   // math.log(0.25), GC.neutral()
    EvalOut r; r.leaf = -1; r.pm = 0.0; r.pr = 0.0;
    if (live) r = eval_placement(M, letter, neutral, 0, L0 + CAL_SPACER, A.cutoff);
    const bool ok = r.leaf >= 0;
    unsigned long long used = 0;                      // letters consumed by the winning configuration
    int n_used = 0;
    if (ok) {
        const int8_t *lo = M.leaf_offsets + r.leaf * L;
        for (int k = 0; k < L; k++) {
            const int o = lo[k];
            if (o >= 0) {
                const int pos = k < L0 ? o : L0 + o;  // index into the sample
                used |= ((bits >> (2 * pos)) & 3ULL) << (2 * n_used);
                n_used++;
            }
        }
        atomicAdd(&s_ok, 1u);
    }
    // letter counts: of the whole sample (weights) and of the letters the winning configuration consumed (scores)
    double n_all[4] = {0.0, 0.0, 0.0, 0.0}, n_use[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < L; k++) {
        const int c = (int)((bits >> (2 * k)) & 3);
#pragma unroll
        for (int q = 0; q < 4; q++) n_all[q] += c == q ? 1.0 : 0.0;
    }
    for (int k = 0; k < n_used; k++) {
        const int c = (int)((used >> (2 * k)) & 3);
#pragma unroll
        for (int q = 0; q < 4; q++) n_use[q] += c == q ? 1.0 : 0.0;
    }
    const double inv_ln2 = 1.0 / A.ln2;
    for (int i = 0; i < A.n_class; i++) {
        const double *lg = cls + 4 * i;
        const double l0 = lg[0], l1 = lg[1], l2 = lg[2], l3 = lg[3];
        // The weight enters the histograms only as a sum of ~1e4-1e6 terms (order-free, tolerance 1e-6): its
        // exponent comes from the letter counts.  The score decides a bucket, which must be the reference's: it is
        // first estimated from the counts (within 1e-12 of the reference's letter-by-letter sum) and recomputed in
        // the reference's order only when the estimate lies within 1e-8 of a bucket boundary.
        const double pw = n_all[0] * l0 + n_all[1] * l1 + n_all[2] * l2 + n_all[3] * l3;
        const double weight = live ? exp(pw + A.K) : 0.0;
        int bin = 0;
        if (ok) {
            const double est = (r.pm - (n_use[0] * l0 + n_use[1] * l1 + n_use[2] * l2 + n_use[3] * l3)) * inv_ln2;
            const double scaled = est * 10.0 + 0.5;
            bin = scaled > 0.0 ? (int)scaled : 0;
            const double frac = scaled - (double)bin;
            if (scaled > -1e-6 && (frac < 1e-7 || frac > 1.0 - 1e-7)) {       // too close to call: the reference's own arithmetic
                double p = 0.0;
                for (int k = 0; k < n_used; k++) p = __dadd_rn(p, lg[(used >> (2 * k)) & 3]);
                double score = __ddiv_rn(__dsub_rn(r.pm, p), A.ln2);
                if (!(score > 0.0)) score = 0.0;
                bin = 0;
                if (score > 0.0) {
                    bin = (int)(score * 10.0 + 0.5);
                    if (bin > A.n_bins) bin = A.n_bins;
                    while (bin > 0 && score < __ldg(A.thr + bin - 1)) bin--;
                    while (bin < A.n_bins && score >= __ldg(A.thr + bin)) bin++;
                }
            }
            if (bin > A.n_bins) bin = A.n_bins;
        }
        double w0 = bin == 0 ? weight : 0.0;          // the common bucket: reduce in the warp first
        for (int o = 16; o; o >>= 1) w0 += __shfl_down_sync(FULL_MASK, w0, o);
        if (lane == 0) bin0[wid][i] += w0;
        if (bin != 0 && live) {
            if (bin >= A.n_bins) atomicOr(A.overflow, 1);
            else atomicAdd(A.hist + (long long)i * A.n_bins + bin, weight);
        }
    }
    __syncthreads();
    for (int i = tid; i < A.n_class; i += CAL_THREADS) {
        double t = 0.0;
        for (int w = 0; w < CAL_THREADS / 32; w++) t += bin0[w][i];
        if (t != 0.0) atomicAdd(A.hist + (long long)i * A.n_bins, t);
    }
    if (tid == 0 && s_ok) atomicAdd(A.n_ok, (unsigned long long)s_ok);
}
