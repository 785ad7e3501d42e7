// table_kernels.cuh — list operations on a device-resident hit table: what rmfilter and rmcluster do
// to a result list, without the list ever becoming Python objects.
//
// Reference semantics restated (paths relative to the reference root):
//   lib/candidates.py:78-97     Candidates.filter      strict thresholds                      -> table_filter_kernel
//   lib/candidates.py:380-394   Candidate.__cmp__      key asc, E-value asc, score desc       -> radix sort on table_key()
//   lib/candidates.py:99-154    top_model / top_seq / top_seq_model  first N of a group in sorted order -> table_group_rank_kernel
//   lib/candidates.py:156-169   get_model                                                     -> table_filter_kernel (model)
//   lib/candidates.py:171-207   nooverlap              order-dependent pairwise elimination   -> nooverlap_* kernels
//   lib/cluster.py:185-203      ClusterAlgorithm.go_cluster, build_index                      -> cluster_* kernels
//   lib/cluster.py:58-83,98-166 Cluster.calc_metrics   counts behind occur / score / bpp / MI / H -> cluster_metrics_kernel
//
// Sorting is a stable LSD radix sort written here (8-bit digits, digits that are equal in every key are skipped);
// Python's list.sort is stable too, so ties keep the table's order exactly as the reference's lists do.
#pragma once
#include "device_common.cuh"

#define RDX_THREADS 256
#define RDX_ITEMS 8
#define RDX_TILE (RDX_THREADS * RDX_ITEMS)      // 2048 keys per block; a warp owns 256 consecutive keys
#define RDX_WARPS (RDX_THREADS / 32)

// key fields a table can be ordered by
enum TableField {
    TF_NEG_SCORE = 0,       // score descending
    TF_EVALUE = 1,          // E-value ascending
    TF_KEY_RANK = 2,        // rank of the sequence name among all names (host-supplied: strings never reach the device)
    TF_MODEL = 3,
    TF_KEY_RANK_MODEL = 4,  // (name, model): top_seq_model's groups
    TF_KEY_FIRST0 = 5,      // (sequence id, first position of chain 0): nooverlap's sweep order
    TF_AUX32 = 6,           // aux[row]
    TF_CL_COL = 7,          // cluster column (chain 1), aux64[2 * row + 1]
    TF_CL_MODEL_ROW = 8     // (model, cluster row (chain 0)), aux64[2 * row]
};

// IEEE double -> unsigned key with the same order (-0.0 and 0.0 compare equal in Python: both map to +0.0)
__device__ __forceinline__ unsigned long long order_bits(double x) {
    if (x == 0.0) x = 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ unsigned long long table_key(const rmd_hit_row &r, const int field, const int *key_rank,
                                                        const uint32_t *aux32, const unsigned long long *aux64, const long long row) {
    switch (field) {
    case TF_NEG_SCORE: return order_bits(-r.score);
    case TF_EVALUE: return order_bits(r.evalue);
    case TF_KEY_RANK: return (unsigned long long)(uint32_t)key_rank[r.key];
    case TF_MODEL: return r.model;
    case TF_KEY_RANK_MODEL: return ((unsigned long long)(uint32_t)key_rank[r.key] << 8) | r.model;
    case TF_KEY_FIRST0: return ((unsigned long long)(uint32_t)r.key << 40) | (unsigned long long)(r.start + r.p0);
    case TF_AUX32: return aux32[row];
    case TF_CL_COL: return aux64[2 * row + 1];
    default: return ((unsigned long long)r.model << 56) | aux64[2 * row];
    }
}

__global__ void table_iota_kernel(uint32_t *idx, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}

__global__ void table_fill_kernel(uint32_t *v, long long n, uint32_t x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = x;
}

// keys[i] = field of row idx[i]
__global__ void table_keys_kernel(const rmd_hit_row *rows, const uint32_t *idx, long long n, int field, const int *key_rank,
                                  const uint32_t *aux32, const unsigned long long *aux64, unsigned long long *keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = idx[i];
    keys[i] = table_key(rows[r], field, key_rank, aux32, aux64, r);
}

// digit totals of all eight bytes of every key: a byte that is the same everywhere needs no pass
__global__ void __launch_bounds__(RDX_THREADS) radix_totals_kernel(const unsigned long long *keys, long long n, unsigned int *totals /* [8][256] */) {
    __shared__ unsigned int h[8 * 256];
    for (int k = threadIdx.x; k < 8 * 256; k += RDX_THREADS) h[k] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * RDX_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RDX_THREADS) {
        const unsigned long long k = keys[i];
#pragma unroll
        for (int b = 0; b < 8; b++) atomicAdd(&h[b * 256 + (int)((k >> (8 * b)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 8 * 256; k += RDX_THREADS) if (h[k]) atomicAdd(&totals[k], h[k]);
}

// per-block digit counts, digit-major: hist[d * nblocks + block]
__global__ void __launch_bounds__(RDX_THREADS) radix_hist_kernel(const unsigned long long *keys, long long n, int shift, unsigned int *hist, int nblocks) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RDX_TILE;
#pragma unroll
    for (int k = 0; k < RDX_ITEMS; k++) {
        const long long i = base + k * RDX_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    hist[(long long)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: offs = exclusive scan of hist.  A warp ranks its 256 consecutive keys 32 at a time (match_any gives the
// lanes with the same digit; the running count per digit lives in shared memory), then the warps' counts are chained.
__global__ void __launch_bounds__(RDX_THREADS) radix_scatter_kernel(const unsigned long long *keys, const uint32_t *idx, long long n, int shift,
                                                                    const unsigned int *offs, int nblocks,
                                                                    unsigned long long *keys_out, uint32_t *idx_out) {
    __shared__ unsigned int wc[RDX_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RDX_WARPS * 256; k += RDX_THREADS) (&wc[0][0])[k] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RDX_TILE + (long long)w * (32 * RDX_ITEMS);
    unsigned long long kk[RDX_ITEMS];
    uint32_t ii[RDX_ITEMS];
    unsigned int rank[RDX_ITEMS];
#pragma unroll
    for (int r = 0; r < RDX_ITEMS; r++) {
        const long long i = base + r * 32 + lane;
        const bool act = i < n;
        kk[r] = act ? keys[i] : 0; ii[r] = act ? idx[i] : 0;
        const int d = act ? (int)((kk[r] >> shift) & 0xff) : 256;
        const unsigned int m = __match_any_sync(FULL_MASK, d);
        rank[r] = 0;
        if (act) {
            rank[r] = wc[w][d] + __popc(m & ((1u << lane) - 1));
            __syncwarp(m);
            if (lane == __ffs(m) - 1) wc[w][d] += __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        unsigned int run = offs[(long long)d * nblocks + blockIdx.x];
        for (int q = 0; q < RDX_WARPS; q++) { const unsigned int t = wc[q][d]; wc[q][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RDX_ITEMS; r++) {
        const long long i = base + r * 32 + lane;
        if (i < n) {
            const unsigned int pos = wc[w][(int)((kk[r] >> shift) & 0xff)] + rank[r];
            keys_out[pos] = kk[r]; idx_out[pos] = ii[r];
        }
    }
}

__global__ void table_gather_kernel(const rmd_hit_row *rows, const uint32_t *idx, long long n, rmd_hit_row *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(rows + idx[i]);
    uint4 *dst = reinterpret_cast<uint4 *>(out + i);
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = src[k];
}

// rows whose flag is set move to out[off[i]] (order kept)
__global__ void table_compact_kernel(const rmd_hit_row *rows, long long n, const uint32_t *flag, const uint32_t *off, rmd_hit_row *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(rows + i);
    uint4 *dst = reinterpret_cast<uint4 *>(out + off[i]);
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = src[k];
}

// is the table in Candidate.__cmp__ order?  first_bad = smallest row whose predecessor compares greater (n: none)
__global__ void table_sorted_kernel(const rmd_hit_row *rows, long long n, const int *key_rank, unsigned long long *first_bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1 || i >= n) return;
    const rmd_hit_row &a = rows[i - 1], &b = rows[i];
    const unsigned long long ka = (unsigned long long)(uint32_t)key_rank[a.key], kb = (unsigned long long)(uint32_t)key_rank[b.key];
    bool greater = ka > kb;
    if (ka == kb) {
        const unsigned long long ea = order_bits(a.evalue), eb = order_bits(b.evalue);
        greater = ea > eb || (ea == eb && order_bits(-a.score) > order_bits(-b.score));
    }
    if (greater) atomicMin(first_bad, (unsigned long long)i);
}

// ---- the last scan's hits become table rows ------------------------------------------------------------------
// letters of the winning gap configuration in node order (chain 0, then chain 1), 3 bits each: 0..3 ACGU, 4 '.', 5 other
__global__ void table_append_scan_kernel(const DevJob J, const rmd_hit *hits, long long n, int key_base, long long start_base,
                                         uint32_t id_base, rmd_hit_row *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rmd_hit h = hits[i];
    const CModel &M = c_models[h.model];
    rmd_hit_row r;
    r.start = h.start + start_base;
    r.key = key_base + h.seq;
    r.wlen = (uint16_t)h.wlen; r.p0 = (uint16_t)h.p0; r.p1 = (uint16_t)h.p1; r.leaf = (uint16_t)h.leaf;
    r.model = (uint8_t)h.model; r.flags = 0; r.pad_ = 0;
    r.score = h.score; r.evalue = h.evalue;
    r.bpp = __longlong_as_double(0x7ff8000000000000ll);          // not computed yet (the reference's None)
    r.letters[0] = r.letters[1] = r.letters[2] = 0;
    const int L0 = M.chain_len[0], L = L0 + M.chain_len[1];
    const long long gbase = J.seq_off[h.seq] + h.start;
    for (int k = 0; k < L && k < 32; k++) {
        const int o = M.leaf_offsets[(long long)h.leaf * L + k];
        uint32_t c = 4;
        if (o >= 0) { const int code = get_code(J, gbase + (k < L0 ? h.p0 : h.p1) + o); c = code < 4 ? (uint32_t)code : 5u; }
        const int bit = 3 * k;
        r.letters[bit >> 5] |= c << (bit & 31);
        if ((bit & 31) > 29) r.letters[(bit >> 5) + 1] |= c >> (32 - (bit & 31));
    }
    r.id = id_base + (uint32_t)i;
    out[i] = r;
}

// ---- Candidates.filter's thresholds / get_model ------------------------------------------------------------------
struct TableFilter {
    int use_score, use_bpp, use_evalue, model;        // model < 0: any
    double min_score, min_bpp, evalue_below;
};
// lib/candidates.py:84-91: drop score <= min_score; drop bpp <= min_bpp (a bpp that was never computed is the reference's
// None, which Python 2 orders below every number: dropped); drop evalue == 0 or -log(evalue) <= min_evalue — the host
// turns the latter into evalue >= evalue_below with the smallest double for which libm's log says so (log is monotone).
__global__ void table_filter_kernel(const rmd_hit_row *rows, long long n, const TableFilter F, uint32_t *flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rmd_hit_row &r = rows[i];
    bool keep = true;
    if (F.use_score && r.score <= F.min_score) keep = false;
    if (F.use_bpp && !(r.bpp > F.min_bpp)) keep = false;
    if (F.use_evalue && (r.evalue == 0.0 || !(r.evalue < F.evalue_below))) keep = false;
    if (F.model >= 0 && r.model != F.model) keep = false;
    flag[i] = keep ? 1u : 0u;
}

// position i of the group-sorted order: its rank inside its group decides whether row idx[i] stays (top_*: dict[key] <= N)
__global__ void table_group_rank_kernel(const unsigned long long *gkeys, const uint32_t *idx, long long n, uint32_t top_n, uint32_t *flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long g = gkeys[i];
    long long lo = 0, hi = i;                               // first position of the group
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (gkeys[mid] < g) lo = mid + 1; else hi = mid; }
    flag[idx[i]] = (uint32_t)(i - lo) < top_n ? 1u : 0u;
}

// ---- nooverlap (lib/candidates.py:171-207) -------------------------------------------------------------------------
// Two hits are "duplicates" when their windows overlap, they belong to one sequence and the position ranges of both chains
// intersect.  The reference walks the list in order: hit i meets every later hit j still alive; a duplicate pair loses j
// when i scores strictly higher and has a positive bpp, else it loses i (and i stops).  The outcome depends on the order, so
// it is replayed exactly — but only inside connected groups: duplicates need intersecting chain-0 ranges on one sequence,
// so a sweep over (sequence, first position of chain 0) with a running maximum of the last position cuts the table into
// groups that cannot interact, and one warp replays the reference's loop on each group in table order.
struct ChainSpan { long long f0, l0, f1, l1; int gap_end; };
__device__ __forceinline__ ChainSpan chain_span(const rmd_hit_row &r) {
    const CModel &M = c_models[r.model];
    const int L0 = M.chain_len[0], L = L0 + M.chain_len[1];
    const int o0 = M.leaf_offsets[(long long)r.leaf * L + L0 - 1], o1 = M.leaf_offsets[(long long)r.leaf * L + L - 1];
    ChainSpan s;
    s.f0 = r.start + r.p0; s.f1 = r.start + r.p1;
    // a chain that ends in a gap has None as its last position: Python 2 orders None below every number, so
    // `pi[0] > pj[-1]` / `pi[-1] < pj[0]` make such a hit nobody's duplicate
    s.gap_end = o0 < 0 || o1 < 0;
    s.l0 = s.f0 + (o0 < 0 ? -1 : o0); s.l1 = s.f1 + (o1 < 0 ? -1 : o1);
    return s;
}

// in sweep order: packed (sequence, last position of chain 0) for the running maximum
__global__ void nooverlap_last_kernel(const rmd_hit_row *rows, const uint32_t *idx, long long n, unsigned long long *last) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rmd_hit_row r = rows[idx[i]];
    const ChainSpan s = chain_span(r);
    last[i] = ((unsigned long long)(uint32_t)r.key << 40) | (unsigned long long)(s.l0 < 0 ? 0 : s.l0);
}

// inclusive running maximum of 64-bit values, three launches like the sums above
__global__ void __launch_bounds__(SCAN_BT) maxscan_block_kernel(const unsigned long long *in, long long n, unsigned long long *bmax) {
    __shared__ unsigned long long tmp[SCAN_BT / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
    unsigned long long m = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) if (base + k < n) m = max(m, in[base + k]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0) tmp[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int k = 1; k < SCAN_BT / 32; k++) m = max(m, tmp[k]); bmax[blockIdx.x] = m; }
}
__global__ void maxscan_carry_kernel(unsigned long long *bmax, int nblocks) {     // one thread: exclusive running maximum over blocks
    unsigned long long run = 0;
    for (int b = 0; b < nblocks; b++) { const unsigned long long v = bmax[b]; bmax[b] = run; run = max(run, v); }
}
__global__ void __launch_bounds__(SCAN_BT) maxscan_apply_kernel(const unsigned long long *in, long long n, const unsigned long long *bmax, unsigned long long *out) {
    __shared__ unsigned long long tmp[SCAN_BT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_IPT;
    unsigned long long v[SCAN_IPT], m = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) { v[k] = base + k < n ? in[base + k] : 0; m = max(m, v[k]); }
    unsigned long long inc = m;                               // inclusive maximum over the threads of the warp
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= o) inc = max(inc, t); }
    if (lane == 31) tmp[wid] = inc;
    __syncthreads();
    unsigned long long before = bmax[blockIdx.x];
    for (int q = 0; q < wid; q++) before = max(before, tmp[q]);
    const unsigned long long prev = __shfl_up_sync(FULL_MASK, inc, 1);
    if (lane > 0) before = max(before, prev);
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) if (base + k < n) { before = max(before, v[k]); out[base + k] = before; }
}

// a group starts where the sweep key exceeds everything seen so far (another sequence, or a gap between chain-0 ranges)
__global__ void nooverlap_breaks_kernel(const unsigned long long *sweep, const unsigned long long *runmax, long long n, uint32_t *brk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    brk[i] = (i == 0 || sweep[i] > runmax[i - 1]) ? 1u : 0u;
}

// group number of every row: comp[idx[i]] = (breaks up to and including i) - 1
__global__ void nooverlap_label_kernel(const uint32_t *idx, const uint32_t *brk, const uint32_t *brk_off, long long n, uint32_t *comp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) comp[idx[i]] = brk_off[i] + brk[i] - 1u;
}

// one warp per group; members = rows `mem[lo .. hi)` in table order (gkeys = their group numbers, sorted)
__global__ void __launch_bounds__(256) nooverlap_replay_kernel(const rmd_hit_row *rows, const uint32_t *mem, const unsigned long long *gkeys, long long n,
                                                               uint32_t n_groups, volatile uint32_t *alive) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long g = warp; g < n_groups; g += n_warps) {
        long long lo = 0, hi = n;                             // [first member of group g, first member of group g + 1)
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (gkeys[mid] < (unsigned long long)g) lo = mid + 1; else hi = mid; }
        const long long m0 = lo;
        hi = n;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (gkeys[mid] <= (unsigned long long)g) lo = mid + 1; else hi = mid; }
        const long long m1 = lo;
        if (m1 - m0 < 2) continue;
        for (long long a = m0; a < m1; a++) {
            if (!alive[mem[a]]) continue;                      // warp-uniform: every lane reads the same word
            const rmd_hit_row ra = rows[mem[a]];
            const ChainSpan sa = chain_span(ra);
            const long long ea = ra.start + ra.wlen;
            const bool a_wins_ties = ra.bpp > 0.0;             // `cands[i].bpp > 0` (None / NaN: false)
            bool a_dead = false;
            for (long long j0 = a + 1; j0 < m1 && !a_dead; j0 += 32) {
                const long long j = j0 + lane;
                bool dup = false;
                double sj = 0.0;
                if (j < m1 && alive[mem[j]]) {
                    const rmd_hit_row rj = rows[mem[j]];
                    const ChainSpan sb = chain_span(rj);
                    sj = rj.score;
                    dup = ra.start < rj.start + rj.wlen && rj.start < ea && ra.key == rj.key && !sa.gap_end && !sb.gap_end &&
                          !(sa.f0 > sb.l0 || sa.l0 < sb.f0) && !(sa.f1 > sb.l1 || sa.l1 < sb.f1);
                }
                uint32_t dm = __ballot_sync(FULL_MASK, dup);
                while (dm) {                                    // in list order, as the reference meets them
                    const int b = __ffs(dm) - 1;
                    dm &= dm - 1;
                    const double s_other = __shfl_sync(FULL_MASK, sj, b);
                    if (ra.score > s_other && a_wins_ties) { if (lane == 0) alive[mem[j0 + b]] = 0; }
                    else { if (lane == 0) alive[mem[a]] = 0; a_dead = true; break; }
                }
                __syncwarp();
            }
            __syncwarp();
        }
    }
}

// ---- rmcluster: the first cluster list (lib/cluster.py:185-203) ----------------------------------------------------
// cluster key of a hit = (model, column of the first position of chain 0, of chain 1); columns come from the alignment's
// column index when the sequences were read from one (lib/candidates.py:489-506), else they are the positions themselves
__global__ void cluster_cols_kernel(const rmd_hit_row *rows, long long n, const long long *col_off, const int *cols, unsigned long long *rc /* [n][2] */,
                                    unsigned int *bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rmd_hit_row r = rows[i];
    long long a = r.start + r.p0, b = r.start + r.p1;
    if (cols) {
        const long long o = col_off[r.key], len = col_off[r.key + 1] - o;
        if (a >= len || b >= len) { atomicOr(bad, 1u); a = b = 0; }       // a position beyond the sequence's column index
        else { a = cols[o + a]; b = cols[o + b]; }
    }
    rc[2 * i] = (unsigned long long)a; rc[2 * i + 1] = (unsigned long long)b;
}

// in (model, row, column) order: does position i open a new cluster key?
__global__ void cluster_heads_kernel(const rmd_hit_row *rows, const uint32_t *idx, const unsigned long long *rc, long long n, uint32_t *head) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool h = i == 0;
    if (!h) {
        const uint32_t a = idx[i], b = idx[i - 1];
        h = rows[a].model != rows[b].model || rc[2 * (long long)a] != rc[2 * (long long)b] || rc[2 * (long long)a + 1] != rc[2 * (long long)b + 1];
    }
    head[i] = h ? 1u : 0u;
}

// the sort was stable, so the head of a key is its first row in table order: clusters are numbered by that row
__global__ void cluster_firsts_kernel(const uint32_t *idx, const uint32_t *head, const uint32_t *head_off, long long n, unsigned long long *first_of_group) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && head[i]) first_of_group[head_off[i]] = idx[i];
}
__global__ void cluster_number_kernel(const uint32_t *gidx, long long n_groups, uint32_t *cluster_of_group) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_groups) cluster_of_group[gidx[r]] = (uint32_t)r;
}
struct ClusterInfo { long long row, col; int model, count, first; int pad_; };
__global__ void cluster_assign_kernel(const rmd_hit_row *rows, const uint32_t *idx, const uint32_t *head, const uint32_t *head_off,
                                      const uint32_t *cluster_of_group, const unsigned long long *rc, long long n,
                                      int *cluster_of_row, ClusterInfo *info) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = idx[i];
    const uint32_t c = cluster_of_group[head_off[i] + head[i] - 1u];
    cluster_of_row[r] = (int)c;
    atomicAdd(&info[c].count, 1);
    if (head[i]) { info[c].row = (long long)rc[2 * (long long)r]; info[c].col = (long long)rc[2 * (long long)r + 1]; info[c].model = rows[r].model; info[c].first = (int)r; }
}

// ---- Cluster.calc_metrics (lib/cluster.py:58-83,98-166): the sums and counts; logs are taken on the host ----------------
// One warp per cluster over its member rows `members[off[c] .. off[c + 1])` (the reference's cluster.cands order):
// score and bpp are summed by one lane in that order (the reference's `+=`), letter pairs of the model's PAIRS are counted
// by all lanes: counts[c][pair][x * 4 + y] over ACGU x ACGU (the host keeps the six canonical combinations).
#define CL_MAX_PAIRS RMD_CL_MAX_PAIRS
struct ClusterPairs { int n[RMD_MAX_MODELS]; unsigned char id[RMD_MAX_MODELS][CL_MAX_PAIRS][2]; };
__device__ __forceinline__ uint32_t row_letter(const rmd_hit_row &r, int k) {
    const int bit = 3 * k;
    uint32_t v = r.letters[bit >> 5] >> (bit & 31);
    if ((bit & 31) > 29) v |= r.letters[(bit >> 5) + 1] << (32 - (bit & 31));
    return v & 7u;
}
__global__ void __launch_bounds__(256) cluster_metrics_kernel(const rmd_hit_row *rows, const int *members, const long long *off, int n_clusters,
                                                              const ClusterPairs P, double *sums /* [c][2] */, unsigned int *counts /* [c][CL_MAX_PAIRS][16] */) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp; c < n_clusters; c += n_warps) {
        const long long m0 = off[c], m1 = off[c + 1];
        if (lane == 0) {
            double s = 0.0, b = 0.0;
            for (long long k = m0; k < m1; k++) { const rmd_hit_row &r = rows[members[k]]; s = __dadd_rn(s, r.score); b = __dadd_rn(b, r.bpp); }
            sums[2 * c] = s; sums[2 * c + 1] = b;
        }
        for (long long k = m0 + lane; k < m1; k += 32) {
            const rmd_hit_row r = rows[members[k]];
            const int np = P.n[r.model];
            for (int q = 0; q < np; q++) {
                const uint32_t x = row_letter(r, P.id[r.model][q][0]), y = row_letter(r, P.id[r.model][q][1]);
                if (x < 4 && y < 4) atomicAdd(&counts[(c * CL_MAX_PAIRS + q) * 16 + x * 4 + y], 1u);
            }
        }
    }
}

// distinct sequences of a cluster: members sorted by (cluster, sequence id); every change of the pair counts once
__global__ void cluster_member_keys_kernel(const rmd_hit_row *rows, const int *members, const int *cluster_of_member, long long n, unsigned long long *keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((unsigned long long)(uint32_t)cluster_of_member[i] << 32) | (uint32_t)rows[members[i]].key;
}
__global__ void cluster_distinct_kernel(const unsigned long long *keys, long long n, unsigned int *n_keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (i == 0 || keys[i] != keys[i - 1])) atomicAdd(&n_keys[keys[i] >> 32], 1u);
}
