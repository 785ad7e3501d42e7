/*
 * rmdetect_b200.h — C ABI of the B200 scan path for RMDetect module models.
 *
 * Plain C, plain pointers and sizes; no torch or CUDA types cross this boundary.
 * Every entry point returns 0 on success or a negative RMD_E_* code; the text of
 * the last failure is available from rmd_last_error().  There is no CPU path
 * behind these calls: without a usable CUDA device rmd_create() fails.
 *
 * What each entry point replaces in the reference (paths relative to the
 * reference root; the reference is pure Python, so "binding" means the call a
 * maintainer would redirect — see INTEGRATION.md for the ctypes stubs):
 *
 *   rmd_set_models      lib/model.py:13-53,148-209 + lib/node.py:38-117  (model tables, built by
 *                       the host flattener rmdetect_b200/modelflat.py and copied to the device)
 *   rmd_set_evalues     lib/evalues.py:38-80 table + :25-36 lookup
 *   rmd_upload          lib/scanner.py:73-75 (GC over the whole sequence is supplied by the host
 *                       or by rmd_gc_counts), lib/bpprobs.py:6-15 (matrix store)
 *   rmd_gc_counts       lib/gcx.py:4-13 (letter histogram; the logs are taken on the host)
 *   rmd_scan_resident   lib/scanner.py:118-159 (search: odometer, gate, eval, hit creation),
 *                       :161-193 (pointer_inc), :195-214 (__get_bpp_mean), :216-244
 *                       (__filter_duplicates), lib/node.py:153-277 (eval / get_solution),
 *                       lib/candidates.py:360-366 (score, E-value), :508-514 (absolute positions)
 *   rmd_scan            the same with host buffers: upload + scan + result download
 *   rmd_set_min_score   rmdetect.py:44 + lib/candidates.py:78-97 (score filter of the CLI, before the download)
 *   rmd_eval            lib/node.py:176-277 called directly (rmcalibrate.py:279-281 style)
 *   rmd_calibrate       rmcalibrate.py:35-62,265-295 (sample scoring into weighted histograms)
 *   rmd_calib_*         the same as a session whose histograms stay on the device (rmcalibrate.py:235-298 loop)
 *   rmd_cpt_counts      lib/bayes.py:69-150 (weighted counts behind rmbuild's conditional probability tables)
 *   rmd_table_*         lib/candidates.py:78-207 (Candidates.filter / top_model / top_seq / top_seq_model / get_model /
 *                       nooverlap, Candidate.__cmp__ order) and lib/cluster.py:58-83,98-166,185-203 (first cluster list,
 *                       the sums and counts of Cluster.calc_metrics) on a hit table that stays on the device
 *   rmd_parse_dot_ps    lib/rnatools.py:176-205 (parse_bpprobs: dot.ps text -> base-pair probabilities)
 *   rmd_compact_bpp     no reference counterpart: lossless 6-byte transport form of a bpp entry
 *   rmd_synth_strand    lib/sequences.py:138-143 (reverse complement), applied to the synthetic stream
 *   rmd_synth_*         no reference counterpart: synthetic genome / stand-in bpp generators,
 *                       bit-identical twins of rmdetect_b200/synth.py, for benchmarks
 */
#ifndef RMDETECT_B200_H
#define RMDETECT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMD_OK 0
#define RMD_E_INVALID (-1)   /* bad argument */
#define RMD_E_CUDA (-2)      /* CUDA runtime failure (text in rmd_last_error) */
#define RMD_E_NOMEM (-3)     /* host or device allocation failed */
#define RMD_E_STATE (-4)     /* call order: models / upload missing */
#define RMD_E_LIMIT (-5)     /* a model or input exceeds a compiled limit */

#define RMD_MAX_MODELS 16
#define RMD_N_CODES 16       /* letter codes: 0-3 ACGU, 4 gap, 5..15 other letters of a sequence */
#define RMD_NSYM 6           /* CPT alphabet: ACGU, gap, other */
#define RMD_MAX_GATE_PAIRS 256
#define RMD_MAX_GATE_CFGS 64
#define RMD_MAX_FIRST_PAIRS 16
#define RMD_MAX_BRANCH_DEPTH 8
#define RMD_FAST_WINDOW 160  /* windows up to this length take the shared-memory kernel */
#define RMD_BIG_WINDOW 256   /* windows of 161..256 letters take the same kernel staged with a larger bit matrix (fewer teams per SM) */

typedef struct rmd_ctx rmd_ctx;

/* search-tree node, DFS preorder (rmdetect_b200/modelflat.py: TREE_DTYPE) */
typedef struct rmd_tree_node {
    int8_t chain, oft;              /* oft < 0: gap at this position in this configuration */
    int8_t par_chain[2], par_oft[2];/* parents on the path; par_oft < 0: the parent is a gap */
    uint8_t flags;                  /* RMD_F_* */
    uint8_t bdepth;                 /* save slot (F_SAVE) / restore slot + 1 (F_RESTORE) */
    uint16_t leaf;                  /* gap configuration id (leaf and distance-check nodes) */
    uint16_t skip;                  /* index of the first node after this subtree */
    uint16_t cpt_row;               /* first CPT row of the model node */
    uint8_t npar, pad_;             /* pad_: left 0 by the caller (the library keeps the node's tree level there) */
} rmd_tree_node;
#define RMD_F_LEAF 1
#define RMD_F_CHECK_DIST 2
#define RMD_F_SAVE 4
#define RMD_F_RESTORE 8

typedef struct rmd_gate_pair { int8_t c1, o1, c2, o2; } rmd_gate_pair;

typedef struct rmd_model_desc {
    int32_t chain_len[2];
    int32_t symmetric;
    int32_t n_tree, n_leaves, n_cpt_rows;
    int32_t n_gate_cfg, n_gate_pairs, n_first;
    int32_t brute_candidates;           /* 1: every placement is a gate candidate */
    const rmd_tree_node *tree;          /* [n_tree] */
    const int32_t *leaf_dist;           /* [n_leaves][2] */
    const double *cpt;                  /* [n_cpt_rows][6], +inf marks a GC row */
    const rmd_gate_pair *gate_pairs;    /* [n_gate_pairs] MUST pairs, configurations concatenated */
    const int32_t *gate_cfg;            /* [n_gate_cfg + 1] */
    const int8_t *first_pairs;          /* [n_first][2] (offset on chain 0, offset on chain 1) */
    const int8_t *leaf_offsets;         /* [n_leaves][chain_len[0] + chain_len[1]] offset within its chain, -1: gap */
} rmd_model_desc;

/* one scan job, host side */
typedef struct rmd_scan_input {
    int32_t n_seq;
    const int64_t *seq_off;     /* [n_seq + 1] letter offsets into codes */
    const uint8_t *codes;       /* one code per letter */
    const double *gc_log;       /* [n_seq][16] log frequency per code (slot 4: log 0.25) */
    const int32_t *gc_class;    /* [n_seq][n_models] selected E-value class, -1: none */
    int32_t win_len, win_step;  /* either 0: each sequence is one window */
    int64_t n_win;              /* windows of all sequences, sequence-major */
    const int64_t *bpp_off;     /* [n_win + 1] entry offsets per window */
    const uint16_t *bpp_i;      /* upper triangle, window-relative, sorted by (i, j) */
    const uint16_t *bpp_j;
    const double *bpp_p;
    double min_bpp_mean;
    double cutoff;              /* natural log, e.g. log(0.1) */
    /* compact transport of the same lists (optional; when bpp_q is given, bpp_i / bpp_j / bpp_p are not read):
       6 bytes per entry instead of 12, see rmd_compact_bpp() */
    const uint16_t *bpp_ij;     /* i | j << 8 (windows up to 256 letters) */
    const uint32_t *bpp_q;      /* round(sqrt(p) * 1e9) | (ulp correction + 1) << 30 */
} rmd_scan_input;

/* one hit, final order = window, model (as given to rmd_set_models), p0, p1 */
typedef struct rmd_hit {
    int64_t window;             /* index into the job's window list */
    int64_t start;              /* window start within its sequence */
    int32_t seq;
    int32_t wlen;
    int32_t p0, p1;             /* window-relative chain pointers */
    int16_t model;
    int16_t leaf;               /* winning gap configuration (index into the model's OFFSETS) */
    int32_t keep;               /* 0: dropped by the duplicate rule */
    double prob_model, prob_rand, score, evalue;
} rmd_hit;

/* The same hit in 32 bytes, for the download (rmd_set_hit_format): what the command line keeps of a hit.  start and
   wlen follow from `window` and the job's geometry (lib/scanner.py:80-111); prob_model / prob_rand are not carried. */
typedef struct rmd_hit32 {
    uint32_t window;            /* index into the job's window list */
    uint32_t seq;
    uint16_t p0, p1;            /* window-relative chain pointers */
    uint8_t model, keep;
    uint16_t leaf;
    double score, evalue;
} rmd_hit32;

typedef struct rmd_scan_result {
    int64_t n_hits;             /* after duplicate removal */
    const void *hits;           /* rmd_hit[n_hits], or rmd_hit32[n_hits] when hit_bytes == 32; host memory owned by the
                                   context, valid until the next scan call */
    int64_t n_raw;              /* before duplicate removal */
    int64_t tries[2];           /* tests_all, tests_pairs */
    const uint32_t *gate_pass;  /* [n_win][n_models] tests_pairs per window and model */
    const uint32_t *raw_count;  /* [n_win][n_models] hits before duplicate removal */
    float ms_scan;              /* device time of the scan kernel(s), CUDA events */
    float ms_post;              /* device time of ordering, scoring and duplicate removal */
    float ms_h2d, ms_d2h;
    float ms_total;             /* device time of the whole call, first enqueue to last copy (CUDA events) */
    int32_t n_launches;         /* kernels launched by this call */
    int32_t hit_bytes;          /* sizeof(rmd_hit) or sizeof(rmd_hit32) */
} rmd_scan_result;

int rmd_create(int device, rmd_ctx **out);
void rmd_destroy(rmd_ctx *ctx);
const char *rmd_last_error(rmd_ctx *ctx);        /* ctx may be NULL: error of a failed rmd_create */
const char *rmd_version(void);

int rmd_set_models(rmd_ctx *ctx, const rmd_model_desc *models, int n_models);
/* table: [n_class][n_rows] E-values, row k = score bucket k/10; thresholds: [n_rows] smallest double
   whose round(x, 1) reaches (k+1)/10 (rmdetect_b200/evalues.py: round_thresholds). n_class 0: no table. */
int rmd_set_evalues(rmd_ctx *ctx, int model, const double *table, int n_class, int n_rows,
                    const double *thresholds);

/* stage a job on the device (sequence packed to 2 bits per letter, bpp coordinate lists) */
int rmd_upload(rmd_ctx *ctx, const rmd_scan_input *in);
/* letter histogram of every uploaded sequence: counts[n_seq][16] */
int rmd_gc_counts(rmd_ctx *ctx, int64_t *counts);
/* replace GC tables / classes of the resident job (after rmd_gc_counts) */
int rmd_set_gc(rmd_ctx *ctx, const double *gc_log, const int32_t *gc_class);
/* 1: rmd_scan* also return the hits removed by the duplicate rule (keep = 0, n_hits = n_raw), which the
   reference shows to its per-window callbacks before lib/scanner.py:53 filters them; default 0 */
int rmd_set_return_dropped(rmd_ctx *ctx, int enable);
/* enable != 0: hits with score <= min_score are dropped on the device, after the duplicate rule, before the download
   (keep = 0 when rmd_set_return_dropped is on; tries and the per-window counters are not affected).  This is the
   command line's own filter (rmdetect.py:44 -> lib/candidates.py:78-97, --min-score, default 5.0) moved in front of
   the copy; the Scanner drop-in leaves it off because the reference's callbacks see unfiltered candidates. */
int rmd_set_min_score(rmd_ctx *ctx, int enable, double min_score);
/* compact != 0: scans return rmd_hit32 records (less than half the bytes on the way back); default 0 */
int rmd_set_hit_format(rmd_ctx *ctx, int compact);
/* Prefix screens (an optimisation with no effect on results): for sequences of at least min_windows windows the
   library tabulates, per sequence, which letter combinations let the first levels of each search tree survive
   (lib/node.py:176-248 evaluated for every combination with the sequence's GC table), and skips the tree walk of
   gate survivors that cannot get past them.  Default 1024; 0: every sequence; < 0: never. */
int rmd_set_screen_policy(rmd_ctx *ctx, int64_t min_windows);
/* scan windows [w_begin, w_end) of the resident job */
int rmd_scan_resident(rmd_ctx *ctx, int64_t w_begin, int64_t w_end, rmd_scan_result *out);
/* upload + scan of all windows + download, host buffers in, host buffers out.  With the compact transport (bpp_q) the
   scan is one launch that runs while the copy stream is still delivering the lists; where the copies cannot run beside a
   launch (a profiler or debugger that serialises the GPU's work) the launch sees no list arrive for 0.2 s, ends, and the
   call scans the resident lists with a plain launch — same results, never an error for that reason. */
int rmd_scan(rmd_ctx *ctx, const rmd_scan_input *in, rmd_scan_result *out);

/* score n placements: sequence k is codes[seq_off[k] .. seq_off[k+1]) with table gc_log[k][16];
   leaf[k] = -1 when eval fails */
int rmd_eval(rmd_ctx *ctx, int model, const uint8_t *codes, const int64_t *seq_off, const double *gc_log,
             const int32_t *p0, const int32_t *p1, int64_t n, double cutoff,
             int32_t *leaf, double *prob_model, double *prob_rand);

/* calibration: samples[n][L] codes 0..3, class_log[n_class][4] = log(gc_i[letter]) in ACGU order (finite: every
   frequency positive, as the reference's math.log demands), hist[n_class][n_bins] is accumulated into (+=).
   Returns via n_ok the samples whose eval held. */
int rmd_calibrate(rmd_ctx *ctx, int model, const uint8_t *samples, int64_t n, int32_t L,
                  const double *class_log, int32_t n_class, double cutoff,
                  double *hist, int32_t n_bins, int64_t *n_ok, int32_t *overflow, float *ms_kernel);
/* same with samples drawn on the device: sample s = the low 2L bits of mix64(seed + GOLDEN * (s + 1)), letter j its
   bits 2j and 2j + 1, s in [first, first + n) (host twin: rmdetect_b200/synth.py synth_calibration_samples) */
int rmd_calibrate_synth(rmd_ctx *ctx, int model, uint64_t seed, int64_t first, int64_t n, int32_t L,
                        const double *class_log, int32_t n_class, double cutoff,
                        double *hist, int32_t n_bins, int64_t *n_ok, int32_t *overflow, float *ms_kernel);

/* Calibration session: the weighted histograms [n_class][n_bins] stay on the device between batches (rmcalibrate.py's
   loop adds 10^4 .. 10^6 samples between two looks at the table).  rmd_calib_open zeroes them; rmd_calib_add enqueues
   one batch — samples[n][L] codes from the host, or, with samples == NULL, drawn on the device as rmd_calibrate_synth
   does — and returns without waiting for it; rmd_calib_device_hist waits for the batches enqueued so far and hands out
   the histogram's device address (n_class * n_bins doubles), so that the GPUs of one job can all-reduce it in place
   (the one collective of the calibration path); rmd_calib_read downloads histogram (may be NULL),
   the number of samples whose eval held, the overflow flag and the device time of the batches since the last read.
   rmd_calibrate / rmd_calibrate_synth end an open session (they are sessions of one batch); the session's buffers are
   its own, other calls on the context leave it alone. */
int rmd_calib_open(rmd_ctx *ctx, int model, int32_t L, const double *class_log, int32_t n_class, double cutoff, int32_t n_bins);
int rmd_calib_add(rmd_ctx *ctx, const uint8_t *samples, uint64_t seed, int64_t first, int64_t n);
int rmd_calib_device_hist(rmd_ctx *ctx, void **dptr, int64_t *n_doubles);
int rmd_calib_read(rmd_ctx *ctx, double *hist, int64_t *n_ok, int32_t *overflow, float *ms_device);
/* rows [first_class, first_class + n_rows) of the session's histogram after the batches enqueued so far — the stop rule of
   rmcalibrate.py:241-257 reads one class (the middle one), so a checkpoint that does not save the table moves 4 KB
   instead of 676 KB; `dptr` (may be NULL) receives the device address of the first of those rows; `cols` (may be NULL)
   the first and the last bucket that is occupied in any class (-1, -1: none) — the score rows of the table run from one
   to the other (rmcalibrate.py:65-83) */
int rmd_calib_read_rows(rmd_ctx *ctx, int32_t first_class, int32_t n_rows, double *rows, void **dptr, int32_t *cols);

/* rmbuild's CPT estimation (lib/bayes.py:69-150 JointProb counting, driven by lib/model.py:100-137 from rmbuild.py:343-352):
   rows of an alignment, one letter code (< K) per model node and row.  For every table t = (node column, parent column 0,
   parent column 1; -1: none) and every row r: cell = l[a] + K * (l[b0] + K * l[b1]) (absent parents count as 0);
   counts[run[r]][t][cell] += 1 and first_row[t][cell] = min(.., r).  `run` numbers the stretches of rows that share one
   weight (rows of one alignment), so that the caller can replay the reference's row-by-row float sums exactly; first_row
   gives the order in which the reference's dicts met their keys.  counts: [n_runs][n_tables][K*K*K] zeroed by the
   library; first_row: [n_tables][K*K*K], INT32_MAX where a cell never occurs. */
int rmd_cpt_counts(rmd_ctx *ctx, const uint8_t *letters, int64_t n_rows, int32_t n_cols, int32_t K,
                   const int32_t *tables /* [n_tables][3] */, int32_t n_tables, const int32_t *run /* [n_rows] */, int32_t n_runs,
                   uint32_t *counts, int32_t *first_row);

/* ---- Device-resident hit table: what rmfilter and rmcluster do to a result list, so that genome-scale
   hit lists never become Python objects.  One table per context.  Rows keep their order unless a call says otherwise;
   every call that removes rows keeps the survivors in order, as the reference's list comprehensions do. */
typedef struct rmd_hit_row {    /* 64 bytes */
    int64_t start;              /* coords[0]: window start within the sequence; coords[1] = start + wlen */
    int32_t key;                /* caller's sequence id: equal ids <=> equal `key` strings */
    uint16_t wlen;
    uint16_t p0, p1;            /* window-relative chain pointers */
    uint16_t leaf;              /* gap configuration */
    uint8_t model, flags;
    uint16_t pad_;
    double score, evalue;
    double bpp;                 /* NaN: not computed (the reference's None) */
    uint32_t letters[3];        /* chain letters in node order (chain 0, then chain 1), 3 bits each: 0..3 ACGU, 4 '.', 5 other */
    uint32_t id;                /* caller's row id, carried through every operation */
} rmd_hit_row;

int rmd_table_clear(rmd_ctx *ctx);
int rmd_table_rows(rmd_ctx *ctx, int64_t *n_rows);
/* append the hits the last rmd_scan / rmd_scan_resident kept (still on the device): key = key_base + seq,
   start = start_base + window start (a piece of a longer sequence), letters read from the resident sequence,
   id = rows appended so far.  first_n >= 0: only the first first_n hits (hits are in window order: a piece's halo
   windows, scanned for the duplicate rule alone, come last) */
int rmd_table_append_scan(rmd_ctx *ctx, int32_t key_base, int64_t start_base, int64_t first_n, int64_t *n_rows);
/* append rows from the host (hits loaded from a .res file) */
int rmd_table_append_rows(rmd_ctx *ctx, const rmd_hit_row *rows, int64_t n, int64_t *n_rows);
int rmd_table_download(rmd_ctx *ctx, rmd_hit_row *rows, int64_t first, int64_t n);
/* lib/candidates.py:78-97 Candidates.filter without its closing sort (rmd_table_sort), and :156-169 get_model: a row stays when
   score > *min_score, bpp > *min_bpp (a bpp that was never computed fails), evalue != 0 and evalue < *evalue_below — the
   caller turns `-math.log(evalue) <= min_evalue` into that bound — and model == `model`; NULL / -1: no such test */
int rmd_table_filter(rmd_ctx *ctx, const double *min_score, const double *min_bpp, const double *evalue_below, int32_t model, int64_t *n_rows);
/* lib/candidates.py:380-394 Candidate.__cmp__ order, stable: key_rank[key] ascending (rank of the key string among all key
   strings, equal strings equal ranks), E-value ascending, score descending */
int rmd_table_sort(rmd_ctx *ctx, const int32_t *key_rank, int32_t n_keys);
/* *first_bad = the first row that compares smaller than its predecessor in that order, -1 when the table is sorted */
int rmd_table_is_sorted(rmd_ctx *ctx, const int32_t *key_rank, int32_t n_keys, int64_t *first_bad);
/* lib/candidates.py:99-154 top_model (group 0) / top_seq (1) / top_seq_model (2): sort, then the first top_n rows of every group */
int rmd_table_top(rmd_ctx *ctx, int32_t group, int32_t top_n, const int32_t *key_rank, int32_t n_keys, int64_t *n_rows);
/* lib/candidates.py:171-207 nooverlap, replayed in table order (the result depends on it, as the reference's does) */
int rmd_table_nooverlap(rmd_ctx *ctx, int64_t *n_rows);
/* lib/cluster.py:185-203: the first cluster list.  Cluster key of a row = (model, column of the first position of chain 0,
   column of the first position of chain 1); column = cols[col_off[key] + position] (the alignment's column index,
   lib/candidates.py:489-506) or the position itself when cols is NULL.  Clusters are numbered in the order the table meets
   their keys; cluster_of[n_rows] receives every row's cluster. */
int rmd_table_cluster_index(rmd_ctx *ctx, const int64_t *col_off, const int32_t *cols, int32_t n_keys, int32_t *cluster_of, int64_t *n_clusters);
/* per cluster of the last rmd_table_cluster_index: its key (row, col, model), member count and first row; any pointer may be NULL */
int rmd_table_cluster_info(rmd_ctx *ctx, int64_t *row, int64_t *col, int32_t *model, int32_t *count, int32_t *first);
/* lib/cluster.py:58-83,98-166 Cluster.calc_metrics, the part that reads hits: for cluster c with member rows
   members[off[c] .. off[c + 1]) (in the order of the reference's cluster.cands): sums[c] = (score, bpp) added in that order,
   n_keys[c] = distinct sequences, counts[c][q][x * 4 + y] = members whose letters at the nodes of the model's q-th PAIRS line
   (pair_ids[model][q] = (id1, id2), RMD_CL_MAX_PAIRS per model, n_pairs[model] used) are x and y in ACGU order */
#define RMD_CL_MAX_PAIRS 16
int rmd_table_cluster_metrics(rmd_ctx *ctx, const int32_t *members, const int64_t *off, int32_t n_clusters,
                              const uint8_t *pair_ids, const int32_t *n_pairs, double *sums, uint32_t *counts, uint32_t *n_keys);

/* synthetic job generated on the device: one sequence of n letters (positions first..first+n-1 of
   the stream keyed by seq_seed), windows per win_len/win_step, stand-in bpp keyed by bpp_seed */
int rmd_synth_job(rmd_ctx *ctx, int64_t n, int64_t first, uint64_t seq_seed, uint64_t bpp_seed,
                  int32_t win_len, int32_t win_step, double min_bpp_mean, double cutoff,
                  const double *gc_log, const int32_t *gc_class, int64_t *n_win, int64_t *n_bpp,
                  float *ms_generate);
/* Strand read by the following rmd_synth_job / rmd_synth_counts calls.  total: letters of the whole stream
   (0: unbounded, forward only).  reverse != 0: position k of the strand is the complement of stream position
   total - 1 - k, i.e. the reverse-complement sequence lib/sequences.py:103-111,138-143 appends for --both-strands. */
int rmd_synth_strand(rmd_ctx *ctx, int64_t total, int32_t reverse);
/* letter counts (A, C, G, U) of strand positions [first, first + n); nothing is stored.  Lets a strand that is
   scanned in pieces get the GC table of the whole sequence (lib/scanner.py:74-75, lib/gcx.py:4-13). */
int rmd_synth_counts(rmd_ctx *ctx, int64_t n, int64_t first, uint64_t seq_seed, int64_t *counts);
/* copy the resident job's sequence codes / one window's bpp list back (tests) */
int rmd_download_codes(rmd_ctx *ctx, uint8_t *codes, int64_t n);
int rmd_download_bpp(rmd_ctx *ctx, int64_t window, uint16_t *i, uint16_t *j, double *p, int64_t cap, int64_t *n);
/* the whole resident bpp store: bpp_off[n_win + 1] and, when the pointers are not NULL, the n_bpp entries */
int rmd_download_job(rmd_ctx *ctx, int64_t *bpp_off, uint16_t *i, uint16_t *j, double *p);

/* Compact form of a coordinate list (host only, no device work).  RNAfold prints sqrt(p) with 9 decimals
   (lib/rnatools.py:176-205 squares it with `** 2.0`), so p = (q / 1e9)^2 up to the last bit of libm's pow:
   q = round(sqrt(p) * 1e9) < 2^30 and c in {-1, 0, +1} ulps are stored, and the device rebuilds exactly the
   caller's double.  Returns RMD_E_INVALID (and the index in *first_bad) for an entry that has no such
   form (i or j > 255, p further than one ulp from a squared 9-decimal number); use the fp64 arrays then. */
int rmd_compact_bpp(const uint16_t *i, const uint16_t *j, const double *p, int64_t n,
                    uint16_t *ij, uint32_t *q, int64_t *first_bad);

/* Parse the text of a ViennaRNA dot.ps (replaces lib/rnatools.py:176-205 parse_bpprobs, reading the file
   RNAfold -p really writes): every `i j sqrt(p) ubox` line becomes an entry (0-based, i < j, p = sqrt_p ** 2.0
   as the reference computes it, zeros dropped, a later line for the same pair replaces the earlier one), the
   result sorted by (i, j).  seq receives the /sequence block (NUL-terminated, at most seq_cap - 1 letters),
   q may be NULL.  Returns RMD_E_LIMIT when cap is too small (*n = entries needed). */
int rmd_parse_dot_ps(const char *text, int64_t len, uint16_t *i, uint16_t *j, double *p, uint32_t *q,
                     int64_t cap, int64_t *n, char *seq, int64_t seq_cap);

/* pinned host staging for callers that want full PCIe rate */
void *rmd_host_alloc(int64_t bytes);
void rmd_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
