// rmd_table.cuh — entry points of the device-resident hit table (include/rmdetect_b200.h: rmd_table_*), included by
// rmd_lib.cu after rmd_ctx.  The kernels are in table_kernels.cuh; this file owns buffers, pass planning and launches.
#pragma once

static int table_launches_dummy = 0;
#define TBL_BLOCKS(n) ((unsigned)(((n) + 255) / 256))

static int table_reserve(rmd_ctx *ctx, long long n) {
    HitTable &T = ctx->table;
    const size_t m = (size_t)std::max<long long>(n, 1);
    RESERVE(T.keys, sizeof(unsigned long long) * m);
    RESERVE(T.keys_alt, sizeof(unsigned long long) * m);
    RESERVE(T.idx, sizeof(uint32_t) * m);
    RESERVE(T.idx_alt, sizeof(uint32_t) * m);
    RESERVE(T.flag, sizeof(uint32_t) * m);
    RESERVE(T.off, sizeof(uint32_t) * (m + 1));
    RESERVE(T.aux32, sizeof(uint32_t) * m);
    const size_t nblocks = (m + RDX_TILE - 1) / RDX_TILE;
    RESERVE(T.hist, sizeof(uint32_t) * 256 * nblocks);
    RESERVE(T.hist_off, sizeof(uint32_t) * (256 * nblocks + 1));
    RESERVE(T.totals, sizeof(uint32_t) * 8 * 256);
    return RMD_OK;
}

// room for n rows in both row buffers, contents of `rows` kept
static int table_reserve_rows(rmd_ctx *ctx, long long n) {
    HitTable &T = ctx->table;
    const size_t bytes = sizeof(rmd_hit_row) * (size_t)std::max<long long>(n, 1);
    if (bytes > T.rows.cap) {
        DBuf bigger;
        RESERVE(bigger, bytes + bytes / 2);
        if (T.n > 0) CK(cudaMemcpyAsync(bigger.p, T.rows.p, sizeof(rmd_hit_row) * (size_t)T.n, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (T.rows.p) cudaFree(T.rows.p);
        T.rows = bigger;
    }
    RESERVE(T.rows_alt, T.rows.cap);
    return RMD_OK;
}

// stable LSD radix sort of the (keys, idx) pairs in T.keys / T.idx; bytes that are equal in every key are skipped
static int table_radix_sort(rmd_ctx *ctx, long long n) {
    HitTable &T = ctx->table;
    if (n < 2) return RMD_OK;
    cudaStream_t st = ctx->stream;
    const int nblocks = (int)((n + RDX_TILE - 1) / RDX_TILE);
    CK(cudaMemsetAsync(T.totals.p, 0, sizeof(uint32_t) * 8 * 256, st));
    radix_totals_kernel<<<std::min(nblocks, 148 * 8), RDX_THREADS, 0, st>>>((const unsigned long long *)T.keys.p, n, (unsigned int *)T.totals.p);
    std::vector<uint32_t> totals(8 * 256);
    CK(cudaMemcpyAsync(totals.data(), T.totals.p, sizeof(uint32_t) * 8 * 256, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int b = 0; b < 8; b++) {
        bool uniform = false;
        for (int d = 0; d < 256; d++) if (totals[b * 256 + d] == (uint32_t)n) uniform = true;
        if (uniform) continue;
        radix_hist_kernel<<<nblocks, RDX_THREADS, 0, st>>>((const unsigned long long *)T.keys.p, n, 8 * b, (unsigned int *)T.hist.p, nblocks);
        int r = device_scan(ctx, (const uint32_t *)T.hist.p, 256LL * nblocks, (uint32_t *)T.hist_off.p, &table_launches_dummy);
        if (r) return r;
        radix_scatter_kernel<<<nblocks, RDX_THREADS, 0, st>>>((const unsigned long long *)T.keys.p, (const uint32_t *)T.idx.p, n, 8 * b,
                                                              (const unsigned int *)T.hist_off.p, nblocks,
                                                              (unsigned long long *)T.keys_alt.p, (uint32_t *)T.idx_alt.p);
        std::swap(T.keys, T.keys_alt);
        std::swap(T.idx, T.idx_alt);
    }
    CK(cudaGetLastError());
    return RMD_OK;
}

// order T.idx (already filled) by one field of the rows it names; the sorted keys stay in T.keys
static int table_sort_by(rmd_ctx *ctx, long long n, int field, const int *d_key_rank, const uint32_t *aux32, const unsigned long long *aux64) {
    HitTable &T = ctx->table;
    if (n < 1) return RMD_OK;
    table_keys_kernel<<<TBL_BLOCKS(n), 256, 0, ctx->stream>>>((const rmd_hit_row *)T.rows.p, (const uint32_t *)T.idx.p, n, field, d_key_rank, aux32, aux64,
                                                             (unsigned long long *)T.keys.p);
    return table_radix_sort(ctx, n);
}

static int table_iota(rmd_ctx *ctx, long long n) {
    if (n > 0) table_iota_kernel<<<TBL_BLOCKS(n), 256, 0, ctx->stream>>>((uint32_t *)ctx->table.idx.p, n);
    return RMD_OK;
}

// keep the rows whose T.flag is 1, in order
static int table_compact(rmd_ctx *ctx) {
    HitTable &T = ctx->table;
    const long long n = T.n;
    if (n == 0) return RMD_OK;
    cudaStream_t st = ctx->stream;
    int r = device_scan(ctx, (const uint32_t *)T.flag.p, n, (uint32_t *)T.off.p, &table_launches_dummy);
    if (r) return r;
    table_compact_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>((const rmd_hit_row *)T.rows.p, n, (const uint32_t *)T.flag.p, (const uint32_t *)T.off.p, (rmd_hit_row *)T.rows_alt.p);
    uint32_t kept = 0;
    CK(cudaMemcpyAsync(&kept, (const uint32_t *)T.off.p + n, sizeof kept, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    std::swap(T.rows, T.rows_alt);
    T.n = kept;
    return RMD_OK;
}

// the host's key ranks on the device (T.ranks), checked against the ids the table holds
static int table_key_ranks(rmd_ctx *ctx, const int32_t *key_rank, int32_t n_keys, const char *who) {
    HitTable &T = ctx->table;
    if (!key_rank || n_keys < 1) return fail(ctx, RMD_E_INVALID, "%s: key_rank is required", who);
    if (T.n > 0 && (T.min_key < 0 || T.max_key >= n_keys)) return fail(ctx, RMD_E_INVALID, "%s: the table holds sequence ids %d..%d, key_rank has %d entries", who, T.min_key, T.max_key, n_keys);
    for (int k = 0; k < n_keys; k++) if (key_rank[k] < 0) return fail(ctx, RMD_E_INVALID, "%s: negative rank", who);
    RESERVE(T.ranks, sizeof(int) * (size_t)n_keys);
    CK(cudaMemcpyAsync(T.ranks.p, key_rank, sizeof(int) * (size_t)n_keys, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RMD_OK;
}

// Candidate.__cmp__ order (lib/candidates.py:380-394): LSD over (score descending, E-value, key)
static int table_sort_cmp(rmd_ctx *ctx) {
    HitTable &T = ctx->table;
    const long long n = T.n;
    if (n < 2) return RMD_OK;
    int r;
    if ((r = table_reserve(ctx, n))) return r;
    table_iota(ctx, n);
    const int fields[3] = {TF_NEG_SCORE, TF_EVALUE, TF_KEY_RANK};
    for (int f : fields) if ((r = table_sort_by(ctx, n, f, (const int *)T.ranks.p, nullptr, nullptr))) return r;
    table_gather_kernel<<<TBL_BLOCKS(n), 256, 0, ctx->stream>>>((const rmd_hit_row *)T.rows.p, (const uint32_t *)T.idx.p, n, (rmd_hit_row *)T.rows_alt.p);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    std::swap(T.rows, T.rows_alt);
    return RMD_OK;
}

#define TABLE_ENTER(who)                                                                            \
    if (!ctx) return fail(ctx, RMD_E_INVALID, who ": NULL context");                                \
    std::lock_guard<std::recursive_mutex> lock_(g_dev_mutex[ctx->device]);                          \
    CK(cudaSetDevice(ctx->device));                                                                 \
    HitTable &T = ctx->table;                                                                       \
    (void)T

extern "C" int rmd_table_clear(rmd_ctx *ctx) {
    TABLE_ENTER("rmd_table_clear");
    T.n = 0; T.next_id = 0; T.min_key = 0x7fffffff; T.max_key = -1; T.n_clusters = 0;
    return RMD_OK;
}

extern "C" int rmd_table_rows(rmd_ctx *ctx, int64_t *n_rows) {
    TABLE_ENTER("rmd_table_rows");
    if (n_rows) *n_rows = T.n;
    return RMD_OK;
}

extern "C" int rmd_table_append_scan(rmd_ctx *ctx, int32_t key_base, int64_t start_base, int64_t first_n, int64_t *n_rows) {
    TABLE_ENTER("rmd_table_append_scan");
    if (!ctx->have_job || !ctx->last_hits) return fail(ctx, RMD_E_STATE, "rmd_table_append_scan: no scan result on the device");
    if (key_base < 0) return fail(ctx, RMD_E_INVALID, "rmd_table_append_scan: negative key_base");
    if (first_n > ctx->last_n) return fail(ctx, RMD_E_INVALID, "rmd_table_append_scan: the scan kept %lld hits, %lld asked for", ctx->last_n, (long long)first_n);
    const long long n = first_n < 0 ? ctx->last_n : first_n;
    if (T.n + n >= (1LL << 32)) return fail(ctx, RMD_E_LIMIT, "rmd_table_append_scan: a table holds fewer than 2^32 rows");
    int r;
    if ((r = bind_tables(ctx))) return r;
    if ((r = table_reserve_rows(ctx, T.n + n))) return r;
    if (n > 0) {
        table_append_scan_kernel<<<TBL_BLOCKS(n), 256, 0, ctx->stream>>>(ctx->job, ctx->last_hits, n, key_base, start_base, T.next_id, (rmd_hit_row *)T.rows.p + T.n);
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        T.min_key = std::min(T.min_key, key_base);
        T.max_key = std::max(T.max_key, key_base + ctx->job.n_seq - 1);
    }
    T.n += n; T.next_id += (uint32_t)n;
    if (n_rows) *n_rows = T.n;
    return RMD_OK;
}

extern "C" int rmd_table_append_rows(rmd_ctx *ctx, const rmd_hit_row *rows, int64_t n, int64_t *n_rows) {
    TABLE_ENTER("rmd_table_append_rows");
    if (n < 0 || (n > 0 && !rows)) return fail(ctx, RMD_E_INVALID, "rmd_table_append_rows: bad arguments");
    if (T.n + n >= (1LL << 32)) return fail(ctx, RMD_E_LIMIT, "rmd_table_append_rows: a table holds fewer than 2^32 rows");
    for (int64_t k = 0; k < n; k++) {
        if (rows[k].key < 0 || rows[k].model >= ctx->n_models || rows[k].start < 0)
            return fail(ctx, RMD_E_INVALID, "rmd_table_append_rows: row %lld names sequence %d / model %d / start %lld", (long long)k, rows[k].key, (int)rows[k].model, (long long)rows[k].start);
        if ((int)rows[k].leaf >= ctx->hmodels[rows[k].model].n_leaves) return fail(ctx, RMD_E_INVALID, "rmd_table_append_rows: row %lld names gap configuration %d", (long long)k, (int)rows[k].leaf);
        T.min_key = std::min(T.min_key, rows[k].key);
        T.max_key = std::max(T.max_key, rows[k].key);
    }
    int r;
    if ((r = table_reserve_rows(ctx, T.n + n))) return r;
    if (n > 0) {
        CK(cudaMemcpyAsync((rmd_hit_row *)T.rows.p + T.n, rows, sizeof(rmd_hit_row) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    T.n += n; T.next_id += (uint32_t)n;
    if (n_rows) *n_rows = T.n;
    return RMD_OK;
}

extern "C" int rmd_table_download(rmd_ctx *ctx, rmd_hit_row *rows, int64_t first, int64_t n) {
    TABLE_ENTER("rmd_table_download");
    if (first < 0 || n < 0 || first + n > T.n || (n > 0 && !rows)) return fail(ctx, RMD_E_INVALID, "rmd_table_download: rows [%lld, %lld) of %lld", (long long)first, (long long)(first + n), T.n);
    if (n > 0) CK(cudaMemcpy(rows, (const rmd_hit_row *)T.rows.p + first, sizeof(rmd_hit_row) * (size_t)n, cudaMemcpyDeviceToHost));
    return RMD_OK;
}

extern "C" int rmd_table_filter(rmd_ctx *ctx, const double *min_score, const double *min_bpp, const double *evalue_below, int32_t model, int64_t *n_rows) {
    TABLE_ENTER("rmd_table_filter");
    if (model >= ctx->n_models) return fail(ctx, RMD_E_INVALID, "rmd_table_filter: model %d", model);
    if (T.n > 0) {
        int r;
        if ((r = table_reserve(ctx, T.n))) return r;
        TableFilter F;
        F.use_score = min_score != nullptr; F.use_bpp = min_bpp != nullptr; F.use_evalue = evalue_below != nullptr; F.model = model;
        F.min_score = min_score ? *min_score : 0.0; F.min_bpp = min_bpp ? *min_bpp : 0.0; F.evalue_below = evalue_below ? *evalue_below : 0.0;
        table_filter_kernel<<<TBL_BLOCKS(T.n), 256, 0, ctx->stream>>>((const rmd_hit_row *)T.rows.p, T.n, F, (uint32_t *)T.flag.p);
        if ((r = table_compact(ctx))) return r;
    }
    if (n_rows) *n_rows = T.n;
    return RMD_OK;
}

extern "C" int rmd_table_sort(rmd_ctx *ctx, const int32_t *key_rank, int32_t n_keys) {
    TABLE_ENTER("rmd_table_sort");
    int r;
    if ((r = table_key_ranks(ctx, key_rank, n_keys, "rmd_table_sort"))) return r;
    return table_sort_cmp(ctx);
}

extern "C" int rmd_table_is_sorted(rmd_ctx *ctx, const int32_t *key_rank, int32_t n_keys, int64_t *first_bad) {
    TABLE_ENTER("rmd_table_is_sorted");
    if (!first_bad) return fail(ctx, RMD_E_INVALID, "rmd_table_is_sorted: NULL argument");
    int r;
    if ((r = table_key_ranks(ctx, key_rank, n_keys, "rmd_table_is_sorted"))) return r;
    *first_bad = -1;
    if (T.n < 2) return RMD_OK;
    RESERVE(T.totals, sizeof(uint32_t) * 8 * 256);
    unsigned long long bad = (unsigned long long)T.n;
    CK(cudaMemcpyAsync(T.totals.p, &bad, sizeof bad, cudaMemcpyHostToDevice, ctx->stream));
    table_sorted_kernel<<<TBL_BLOCKS(T.n), 256, 0, ctx->stream>>>((const rmd_hit_row *)T.rows.p, T.n, (const int *)T.ranks.p, (unsigned long long *)T.totals.p);
    CK(cudaMemcpyAsync(&bad, T.totals.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if ((long long)bad < T.n) *first_bad = (int64_t)bad;
    return RMD_OK;
}

extern "C" int rmd_table_top(rmd_ctx *ctx, int32_t group, int32_t top_n, const int32_t *key_rank, int32_t n_keys, int64_t *n_rows) {
    TABLE_ENTER("rmd_table_top");
    if (group < 0 || group > 2 || top_n < 0) return fail(ctx, RMD_E_INVALID, "rmd_table_top: group %d, n %d", group, top_n);
    int r;
    if ((r = table_key_ranks(ctx, key_rank, n_keys, "rmd_table_top"))) return r;
    if ((r = table_sort_cmp(ctx))) return r;                                   // every top_* starts with cands.sort()
    const long long n = T.n;
    if (n > 0) {
        if ((r = table_reserve(ctx, n))) return r;
        table_iota(ctx, n);
        const int field = group == 0 ? TF_MODEL : group == 1 ? TF_KEY_RANK : TF_KEY_RANK_MODEL;
        if ((r = table_sort_by(ctx, n, field, (const int *)T.ranks.p, nullptr, nullptr))) return r;
        table_group_rank_kernel<<<TBL_BLOCKS(n), 256, 0, ctx->stream>>>((const unsigned long long *)T.keys.p, (const uint32_t *)T.idx.p, n, (uint32_t)top_n, (uint32_t *)T.flag.p);
        if ((r = table_compact(ctx))) return r;
    }
    if (n_rows) *n_rows = T.n;
    return RMD_OK;
}

extern "C" int rmd_table_nooverlap(rmd_ctx *ctx, int64_t *n_rows) {
    TABLE_ENTER("rmd_table_nooverlap");
    const long long n = T.n;
    int r;
    if (n > 1) {
        if (T.max_key >= (1 << 24)) return fail(ctx, RMD_E_LIMIT, "rmd_table_nooverlap: sequence ids must stay below 2^24");
        if ((r = bind_tables(ctx))) return r;                                  // chain ends come from the models' gap configurations
        if ((r = table_reserve(ctx, n))) return r;
        RESERVE(T.aux64, sizeof(unsigned long long) * (size_t)n);
        cudaStream_t st = ctx->stream;
        const rmd_hit_row *rows = (const rmd_hit_row *)T.rows.p;
        table_iota(ctx, n);
        if ((r = table_sort_by(ctx, n, TF_KEY_FIRST0, nullptr, nullptr, nullptr))) return r;      // sweep order; sweep keys in T.keys
        nooverlap_last_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>(rows, (const uint32_t *)T.idx.p, n, (unsigned long long *)T.aux64.p);
        const int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
        RESERVE(T.bmax, sizeof(unsigned long long) * (size_t)nb);
        maxscan_block_kernel<<<nb, SCAN_BT, 0, st>>>((const unsigned long long *)T.aux64.p, n, (unsigned long long *)T.bmax.p);
        maxscan_carry_kernel<<<1, 1, 0, st>>>((unsigned long long *)T.bmax.p, nb);
        maxscan_apply_kernel<<<nb, SCAN_BT, 0, st>>>((const unsigned long long *)T.aux64.p, n, (const unsigned long long *)T.bmax.p, (unsigned long long *)T.keys_alt.p);
        nooverlap_breaks_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>((const unsigned long long *)T.keys.p, (const unsigned long long *)T.keys_alt.p, n, (uint32_t *)T.flag.p);
        if ((r = device_scan(ctx, (const uint32_t *)T.flag.p, n, (uint32_t *)T.off.p, &table_launches_dummy))) return r;
        nooverlap_label_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>((const uint32_t *)T.idx.p, (const uint32_t *)T.flag.p, (const uint32_t *)T.off.p, n, (uint32_t *)T.aux32.p);
        uint32_t n_groups = 0;
        CK(cudaMemcpyAsync(&n_groups, (const uint32_t *)T.off.p + n, sizeof n_groups, cudaMemcpyDeviceToHost, st));
        table_iota(ctx, n);
        if ((r = table_sort_by(ctx, n, TF_AUX32, nullptr, (const uint32_t *)T.aux32.p, nullptr))) return r;   // members of a group together, in table order
        table_fill_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>((uint32_t *)T.flag.p, n, 1u);
        const unsigned grid = (unsigned)std::min<long long>(((long long)n_groups + 7) / 8, 148 * 16);
        nooverlap_replay_kernel<<<std::max(grid, 1u), 256, 0, st>>>(rows, (const uint32_t *)T.idx.p, (const unsigned long long *)T.keys.p, n, n_groups, (uint32_t *)T.flag.p);
        if ((r = table_compact(ctx))) return r;
    }
    if (n_rows) *n_rows = T.n;
    return RMD_OK;
}

extern "C" int rmd_table_cluster_index(rmd_ctx *ctx, const int64_t *col_off, const int32_t *cols, int32_t n_keys, int32_t *cluster_of, int64_t *n_clusters) {
    TABLE_ENTER("rmd_table_cluster_index");
    if (!cluster_of || !n_clusters) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_index: NULL argument");
    if ((cols != nullptr) != (col_off != nullptr)) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_index: cols and col_off go together");
    const long long n = T.n;
    *n_clusters = 0; T.n_clusters = 0;
    if (n == 0) return RMD_OK;
    if (cols && (T.min_key < 0 || T.max_key >= n_keys)) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_index: the table holds sequence ids %d..%d, col_off has %d sequences", T.min_key, T.max_key, n_keys);
    int r;
    if ((r = table_reserve(ctx, n))) return r;
    cudaStream_t st = ctx->stream;
    RESERVE(T.aux64, sizeof(unsigned long long) * 2 * (size_t)n);
    const long long *d_off = nullptr; const int *d_cols = nullptr;
    if (cols) {
        const long long total = col_off[n_keys];
        RESERVE(T.col_off, sizeof(long long) * (size_t)(n_keys + 1));
        RESERVE(T.cols, sizeof(int) * (size_t)std::max<long long>(total, 1));
        CK(cudaMemcpyAsync(T.col_off.p, col_off, sizeof(long long) * (size_t)(n_keys + 1), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(T.cols.p, cols, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice, st));
        d_off = (const long long *)T.col_off.p; d_cols = (const int *)T.cols.p;
    }
    const rmd_hit_row *rows = (const rmd_hit_row *)T.rows.p;
    const unsigned long long *rc = (const unsigned long long *)T.aux64.p;
    CK(cudaMemsetAsync(T.totals.p, 0, sizeof(uint32_t), st));
    cluster_cols_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>(rows, n, d_off, d_cols, (unsigned long long *)T.aux64.p, (unsigned int *)T.totals.p);
    {
        unsigned int bad = 0;
        CK(cudaMemcpyAsync(&bad, T.totals.p, sizeof bad, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (bad) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_index: a hit lies beyond its sequence's column index");
    }
    table_iota(ctx, n);
    if ((r = table_sort_by(ctx, n, TF_CL_COL, nullptr, nullptr, rc))) return r;
    if ((r = table_sort_by(ctx, n, TF_CL_MODEL_ROW, nullptr, nullptr, rc))) return r;
    cluster_heads_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>(rows, (const uint32_t *)T.idx.p, rc, n, (uint32_t *)T.flag.p);
    if ((r = device_scan(ctx, (const uint32_t *)T.flag.p, n, (uint32_t *)T.off.p, &table_launches_dummy))) return r;
    uint32_t n_groups = 0;
    CK(cudaMemcpyAsync(&n_groups, (const uint32_t *)T.off.p + n, sizeof n_groups, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(T.aux32.p, T.idx.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, st));     // the (model, row, column) order
    cluster_firsts_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>((const uint32_t *)T.aux32.p, (const uint32_t *)T.flag.p, (const uint32_t *)T.off.p, n, (unsigned long long *)T.keys.p);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    table_iota(ctx, n_groups);
    if ((r = table_radix_sort(ctx, n_groups))) return r;                        // keys ordered by their first row = the order build_index meets them
    RESERVE(T.grp, sizeof(uint32_t) * (size_t)n_groups);
    RESERVE(T.cl_info, sizeof(ClusterInfo) * (size_t)n_groups);
    RESERVE(T.cl_of_row, sizeof(int) * (size_t)n);
    CK(cudaMemsetAsync(T.cl_info.p, 0, sizeof(ClusterInfo) * (size_t)n_groups, st));
    cluster_number_kernel<<<TBL_BLOCKS(n_groups), 256, 0, st>>>((const uint32_t *)T.idx.p, n_groups, (uint32_t *)T.grp.p);
    cluster_assign_kernel<<<TBL_BLOCKS(n), 256, 0, st>>>(rows, (const uint32_t *)T.aux32.p, (const uint32_t *)T.flag.p, (const uint32_t *)T.off.p, (const uint32_t *)T.grp.p,
                                                         rc, n, (int *)T.cl_of_row.p, (ClusterInfo *)T.cl_info.p);
    CK(cudaMemcpyAsync(cluster_of, T.cl_of_row.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    T.n_clusters = n_groups;
    *n_clusters = n_groups;
    return RMD_OK;
}

extern "C" int rmd_table_cluster_info(rmd_ctx *ctx, int64_t *row, int64_t *col, int32_t *model, int32_t *count, int32_t *first) {
    TABLE_ENTER("rmd_table_cluster_info");
    std::vector<ClusterInfo> info((size_t)T.n_clusters);
    if (T.n_clusters) CK(cudaMemcpy(info.data(), T.cl_info.p, sizeof(ClusterInfo) * info.size(), cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < info.size(); c++) {
        if (row) row[c] = info[c].row;
        if (col) col[c] = info[c].col;
        if (model) model[c] = info[c].model;
        if (count) count[c] = info[c].count;
        if (first) first[c] = info[c].first;
    }
    return RMD_OK;
}

extern "C" int rmd_table_cluster_metrics(rmd_ctx *ctx, const int32_t *members, const int64_t *off, int32_t n_clusters,
                                         const uint8_t *pair_ids, const int32_t *n_pairs,
                                         double *sums, uint32_t *counts, uint32_t *n_keys) {
    TABLE_ENTER("rmd_table_cluster_metrics");
    if (n_clusters < 0 || !off || !pair_ids || !n_pairs || !sums || !counts || !n_keys) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_metrics: NULL argument");
    if (n_clusters == 0) return RMD_OK;
    const long long nm = off[n_clusters];
    if (off[0] != 0 || nm < 0 || (nm > 0 && !members)) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_metrics: bad member offsets");
    std::vector<int> cl_of((size_t)std::max<long long>(nm, 1));
    for (int c = 0; c < n_clusters; c++) {
        if (off[c + 1] < off[c]) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_metrics: member offsets are not monotone");
        for (long long k = off[c]; k < off[c + 1]; k++) {
            if (members[k] < 0 || members[k] >= T.n) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_metrics: member %lld names row %d of %lld", k, members[k], T.n);
            cl_of[(size_t)k] = c;
        }
    }
    ClusterPairs P;
    memset(&P, 0, sizeof P);
    for (int m = 0; m < ctx->n_models; m++) {
        if (n_pairs[m] < 0 || n_pairs[m] > CL_MAX_PAIRS) return fail(ctx, RMD_E_LIMIT, "rmd_table_cluster_metrics: model %d has %d pairs (limit %d)", m, n_pairs[m], CL_MAX_PAIRS);
        P.n[m] = n_pairs[m];
        const int L = ctx->hmodels[m].chain_len[0] + ctx->hmodels[m].chain_len[1];
        for (int q = 0; q < n_pairs[m]; q++)
            for (int s = 0; s < 2; s++) {
                const int id = pair_ids[((size_t)m * CL_MAX_PAIRS + q) * 2 + s];
                if (id >= L || id >= 32) return fail(ctx, RMD_E_INVALID, "rmd_table_cluster_metrics: model %d pair %d names node %d", m, q, id);
                P.id[m][q][s] = (unsigned char)id;
            }
    }
    int r;
    if ((r = table_reserve(ctx, std::max<long long>(nm, T.n)))) return r;
    cudaStream_t st = ctx->stream;
    RESERVE(T.members, sizeof(int) * (size_t)std::max<long long>(nm, 1));
    RESERVE(T.cl_of_row, sizeof(int) * (size_t)std::max<long long>(std::max<long long>(nm, T.n), 1));
    RESERVE(T.col_off, sizeof(long long) * (size_t)(n_clusters + 1));
    RESERVE(T.sums, sizeof(double) * 2 * (size_t)n_clusters);
    RESERVE(T.counts, sizeof(uint32_t) * CL_MAX_PAIRS * 16 * (size_t)n_clusters);
    RESERVE(T.grp, sizeof(uint32_t) * (size_t)n_clusters);
    if (nm) CK(cudaMemcpyAsync(T.members.p, members, sizeof(int) * (size_t)nm, cudaMemcpyHostToDevice, st));
    if (nm) CK(cudaMemcpyAsync(T.cl_of_row.p, cl_of.data(), sizeof(int) * (size_t)nm, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(T.col_off.p, off, sizeof(long long) * (size_t)(n_clusters + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(T.counts.p, 0, sizeof(uint32_t) * CL_MAX_PAIRS * 16 * (size_t)n_clusters, st));
    CK(cudaMemsetAsync(T.grp.p, 0, sizeof(uint32_t) * (size_t)n_clusters, st));
    const rmd_hit_row *rows = (const rmd_hit_row *)T.rows.p;
    const unsigned grid = (unsigned)std::min<long long>(((long long)n_clusters + 7) / 8, 148 * 16);
    cluster_metrics_kernel<<<grid, 256, 0, st>>>(rows, (const int *)T.members.p, (const long long *)T.col_off.p, n_clusters, P, (double *)T.sums.p, (unsigned int *)T.counts.p);
    if (nm) {
        cluster_member_keys_kernel<<<TBL_BLOCKS(nm), 256, 0, st>>>(rows, (const int *)T.members.p, (const int *)T.cl_of_row.p, nm, (unsigned long long *)T.keys.p);
        table_iota(ctx, nm);
        if ((r = table_radix_sort(ctx, nm))) return r;
        cluster_distinct_kernel<<<TBL_BLOCKS(nm), 256, 0, st>>>((const unsigned long long *)T.keys.p, nm, (unsigned int *)T.grp.p);
    }
    CK(cudaMemcpyAsync(sums, T.sums.p, sizeof(double) * 2 * (size_t)n_clusters, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(counts, T.counts.p, sizeof(uint32_t) * CL_MAX_PAIRS * 16 * (size_t)n_clusters, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_keys, T.grp.p, sizeof(uint32_t) * (size_t)n_clusters, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return RMD_OK;
}
