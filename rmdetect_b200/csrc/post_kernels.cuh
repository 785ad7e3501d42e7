// post_kernels.cuh — K2 (letter counts / 2-bit packing), K3 (ordering, duplicate removal,
// compaction) and K4 (score + E-value), all on the device so only the final hit list is copied back.
//
// Reference semantics restated (paths relative to the reference root):
//   lib/gcx.py:4-13            GC.compute letter counts            -> count_letters_kernel
//   lib/candidates.py:360-366  score = (pm - pr) / log(2.0)         -> finalize_hits_kernel
//   lib/evalues.py:25-36       key "%.1f" % round(score, 1) lookup  -> evalue_of()
//   lib/candidates.py:508-514  update_chains_pos (window start)     -> rmd_hit.start
//   lib/scanner.py:216-244     __filter_duplicates                  -> dedupe_kernel
#pragma once
#include "device_common.cuh"

// ---- exclusive scan of uint32 over n items (3 launches: block sums, scan of sums, rescan) ------
#define SCAN_BT 512
#define SCAN_IPT 8
#define SCAN_TILE (SCAN_BT * SCAN_IPT)

__global__ void __launch_bounds__(SCAN_BT) scan_block_sums_kernel(const uint32_t *in, long long n, uint32_t *sums) {
    __shared__ uint32_t tmp[SCAN_BT / 32 + 1];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan<SCAN_BT>(s, tmp, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BT) scan_sums_kernel(uint32_t *sums, int nblocks, uint32_t *grand_total) {
    __shared__ uint32_t tmp[SCAN_BT / 32 + 1];
    uint32_t carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += SCAN_BT) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nblocks ? sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan<SCAN_BT>(v, tmp, &total);
        if (i < nblocks) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_BT) scan_apply_kernel(const uint32_t *in, long long n, const uint32_t *sums,
                                                             uint32_t *out) {
    __shared__ uint32_t tmp[SCAN_BT / 32 + 1];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    uint32_t total;
    uint32_t run = sums[blockIdx.x] + block_exclusive_scan<SCAN_BT>(s, tmp, &total);
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) if (base + k < n) { out[base + k] = run; run += v[k]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_BT - 1) out[n] = run;   // out has n + 1 slots
}

// ---- K4: E-value of a score --------------------------------------------------------------------
// bucket(x) = #{k : thr[k] <= |x|} equals the tenths index of Python's round(x, 1); a negative score
// that does not round to -0.0, or a bucket beyond the table, misses the table -> 0.0; no table -> 1.0.
__device__ __forceinline__ double evalue_of(const CModel &M, int cls, double score) {
    if (M.ev_nclass == 0 || cls < 0) return 1.0;
    const double a = fabs(score);
    int lo = 0, hi = M.ev_nrows;                    // count of thresholds <= a
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(M.ev_thr + mid) <= a) lo = mid + 1; else hi = mid;
    }
    if (lo >= M.ev_nrows) return 0.0;
    if (score < 0.0 && lo > 0) return 0.0;
    return __ldg(M.ev_table + (long long)cls * M.ev_nrows + lo);
}

// ---- K3: order raw records (window, model, odometer), fill in score / E-value / coordinates ----
// Records arrive in completion order.  place_hits_kernel buckets them by (window, model) group;
// finalize_hits_kernel ranks each record inside its group by (p0, p1) — the odometer order of
// lib/scanner.py:161-193 — by counting smaller keys (groups hold a handful of hits; the count is
// O(group size) per record), and writes the finished hit at its final position.
__global__ void place_hits_kernel(const RawHit *raw, long long n_raw, int n_models, const uint32_t *group_off,
                                  uint32_t *group_fill, RawHit *placed) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_raw) return;
    const RawHit h = raw[r];
    const long long grp = (long long)h.window * n_models + h.model;
    placed[group_off[grp] + atomicAdd(&group_fill[grp], 1u)] = h;
}

__global__ void finalize_hits_kernel(const DevJob J, long long w_begin, const RawHit *placed, long long n_raw,
                                     const uint32_t *group_off, double ln2, rmd_hit *out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_raw) return;
    const RawHit h = placed[r];
    const long long grp = (long long)h.window * J.n_models + h.model;
    const uint32_t lo = group_off[grp], hi = group_off[grp + 1];
    const uint32_t key = ((uint32_t)h.p0 << 16) | h.p1;
    uint32_t rank = 0;
    for (uint32_t q = lo; q < hi; q++) rank += (((uint32_t)placed[q].p0 << 16) | placed[q].p1) < key;
    const long long dst = (long long)lo + rank;
    int seq, wlen;
    long long start;
    window_geom(J, w_begin + h.window, seq, start, wlen);
    rmd_hit o;
    o.window = w_begin + h.window;
    o.start = start;
    o.seq = seq;
    o.wlen = wlen;
    o.p0 = h.p0; o.p1 = h.p1;
    o.model = (int16_t)h.model; o.leaf = (int16_t)h.leaf;
    o.keep = 1;
    o.prob_model = h.pm; o.prob_rand = h.pr;
    o.score = __ddiv_rn(__dsub_rn(h.pm, h.pr), ln2);
    o.evalue = evalue_of(c_models[h.model], J.gc_class[seq * J.n_models + h.model], o.score);
    out[dst] = o;
}

// lib/scanner.py:216-244: an earlier hit is dropped when a later window of the same sequence, whose
// coordinates overlap, holds a hit of the same model with the same first position in every chain.
// (Scores of such twins are equal — same letters, same GC — so the reference always drops the
// earlier one.)  Hits are ordered, so the twin is found by binary search in the later window's group.
// With `filter_score` the score filter of the command line (rmdetect.py:44, lib/candidates.py:84: drop
// score <= min_score) is applied here as well, so that a genome-scale scan downloads only what the CLI keeps;
// twins have equal scores, so filtering before or after the duplicate rule gives the same list.
__global__ void dedupe_kernel(const DevJob J, long long w_begin, long long w_end, rmd_hit *hits, long long n,
                              const uint32_t *group_off, uint32_t *keep, const int filter_score, const double min_score) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rmd_hit h = hits[i];
    if (filter_score && h.score <= min_score) { keep[i] = 0; hits[i].keep = 0; return; }
    const long long a0 = h.start + h.p0, a1 = h.start + h.p1, e1 = h.start + h.wlen;
    const long long w_last = J.win_base[h.seq + 1];
    uint32_t k = 1;
    for (long long w2 = h.window + 1; w2 < w_last && w2 < w_end; w2++) {
        int seq2, wlen2;
        long long s2;
        window_geom(J, w2, seq2, s2, wlen2);
        if (!(h.start < s2 + wlen2 && s2 < e1)) break;      // windows are ordered by start
        const long long q0 = a0 - s2, q1 = a1 - s2;
        if (q0 < 0 || q1 < 0) continue;
        const long long grp = (w2 - w_begin) * J.n_models + h.model;
        long long lo = group_off[grp], hi = group_off[grp + 1];
        const long long key = (q0 << 20) | q1;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const long long km = ((long long)hits[mid].p0 << 20) | hits[mid].p1;
            if (km < key) lo = mid + 1; else hi = mid;
        }
        if (lo < (long long)group_off[grp + 1] && hits[lo].p0 == q0 && hits[lo].p1 == q1) { k = 0; break; }
    }
    keep[i] = k;
    hits[i].keep = (int32_t)k;
}

__global__ void compact_hits_kernel(const rmd_hit *hits, long long n, const uint32_t *keep, const uint32_t *keep_off,
                                    rmd_hit *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keep[i]) out[keep_off[i]] = hits[i];
}

// the download form of a hit list (rmd_set_hit_format)
__global__ void pack_hits32_kernel(const rmd_hit *hits, long long n, rmd_hit32 *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rmd_hit h = hits[i];
    rmd_hit32 o;
    o.window = (uint32_t)h.window; o.seq = (uint32_t)h.seq;
    o.p0 = (uint16_t)h.p0; o.p1 = (uint16_t)h.p1;
    o.model = (uint8_t)h.model; o.keep = (uint8_t)h.keep; o.leaf = (uint16_t)h.leaf;
    o.score = h.score; o.evalue = h.evalue;
    out[i] = o;
}

// ---- compact bpp transport -> the coordinate lists the scan reads (see rmd_compact_bpp) --------------
// p = (q / 1e9)^2 moved by c - 1 ulps (c = top two bits): exactly the double the caller held.
__global__ void expand_bpp_kernel(const uint16_t *ij, const uint32_t *q, long long n, uint16_t *oi, uint16_t *oj, double *op) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t w = ij[k], v = q[k];
    const double x = transport_root(v & 0x3fffffffu);
    const long long bits = __double_as_longlong(__dmul_rn(x, x)) + (long long)(v >> 30) - 1;
    oi[k] = (uint16_t)(w & 0xffu);
    oj[k] = (uint16_t)(w >> 8);
    op[k] = __longlong_as_double(bits);
}

// ---- K2: pack letter codes to 2 bits, flag non-ACGU letters, count letters per sequence ---------
__global__ void pack_codes_kernel(const uint8_t *codes8, long long n, uint32_t *codes2, int *any_other) {
    const long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one output word = 16 letters
    const long long base = wi * 16;
    if (base >= n) return;
    uint32_t word = 0;
    int other = 0;
    if (base + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(codes8 + base);       // coalesced 16-byte load
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t c = (q[k >> 2] >> (8 * (k & 3))) & 0xff;
            other |= c > 3;
            word |= (c & 3) << (2 * k);
        }
    } else {
        for (int k = 0; base + k < n; k++) {
            const uint32_t c = codes8[base + k];
            other |= c > 3;
            word |= (c & 3) << (2 * k);
        }
    }
    codes2[wi] = word;
    if (other) atomicOr(any_other, 1);
}

__global__ void count_letters_kernel(const DevJob J, unsigned long long *counts /* [n_seq][16] */) {
    __shared__ unsigned int local[RMD_N_CODES];
    const int seq = blockIdx.y;
    if (threadIdx.x < RMD_N_CODES) local[threadIdx.x] = 0;
    __syncthreads();
    const long long s0 = J.seq_off[seq], s1 = J.seq_off[seq + 1];
    unsigned int c4[4] = {0, 0, 0, 0};
    for (long long g = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; g < s1; g += (long long)gridDim.x * blockDim.x) {
        const int c = get_code(J, g);
        if (c < 4) c4[c]++; else atomicAdd(&local[c], 1u);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        unsigned int v = c4[c];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&local[c], v);
    }
    __syncthreads();
    if (threadIdx.x < RMD_N_CODES && local[threadIdx.x])
        atomicAdd(&counts[seq * RMD_N_CODES + threadIdx.x], (unsigned long long)local[threadIdx.x]);
}

// placement scoring on demand (rmd_eval): one thread per (sequence, p0, p1)
__global__ void eval_batch_kernel(int model, const uint8_t *codes, const long long *seq_off, const double *gc_log,
                                  const int *p0, const int *p1, long long n, double cutoff,
                                  int *leaf, double *pm, double *pr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *s = codes + seq_off[i];
    const int len = (int)(seq_off[i + 1] - seq_off[i]);
    const CModel &M = c_models[model];
    double gc[RMD_N_CODES];
#pragma unroll
    for (int k = 0; k < RMD_N_CODES; k++) gc[k] = gc_log[i * RMD_N_CODES + k];
    auto letter = [&](int pos) { return pos < len ? (int)s[pos] : 0; };
    const EvalOut r = eval_placement(M, letter, gc, p0[i], p1[i], cutoff);
    leaf[i] = r.leaf; pm[i] = r.pm; pr[i] = r.pr;
}

// ---- rmbuild: counts behind the conditional probability tables (lib/bayes.py:109-127) --------------------------
// one thread per (row, table); cells are counted per run of equally weighted rows, and the first row of every cell is
// kept (the order in which the reference's dicts meet their keys)
__global__ void cpt_count_kernel(const uint8_t *letters, long long n_rows, int n_cols, int K, const int *tables, int n_tables,
                                 const int *run, unsigned int *counts, int *first_row) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * n_tables) return;
    const long long r = i / n_tables;
    const int t = (int)(i - r * n_tables);
    const uint8_t *row = letters + r * n_cols;
    const int a = tables[3 * t], b0 = tables[3 * t + 1], b1 = tables[3 * t + 2];
    const int cell = row[a] + K * ((b0 >= 0 ? row[b0] : 0) + K * (b1 >= 0 ? row[b1] : 0));
    const int k3 = K * K * K;
    atomicAdd(&counts[((long long)run[r] * n_tables + t) * k3 + cell], 1u);
    atomicMin(&first_row[(long long)t * k3 + cell], (int)r);
}
