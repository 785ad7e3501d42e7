/*
 * rmd_oracle.c — CPU restatement of RMDetect's scan hot path.  TEST INFRASTRUCTURE.
 *
 * This file is the parity oracle for the CUDA path.  It is imported only by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
 * reference legs; the product (rmdetect_b200/) never links, loads or calls it.
 *
 * It follows the reference's algorithm statement by statement, with the same
 * fp64 additions in the same order, from these reference files (all paths
 * relative to the reference root):
 *
 *   lib/node.py:38-117     NodeSearch.__init__   -> build_tree()
 *   lib/node.py:143-151    get_parents           -> resolve_parents()
 *   lib/node.py:153-174    get_solution          -> orc_eval() tail
 *   lib/node.py:176-277    NodeSearch.eval       -> eval_node()
 *   lib/scanner.py:73-113  Scanner.slide         -> orc_windows()
 *   lib/scanner.py:118-159 Scanner.search        -> search_window()
 *   lib/scanner.py:161-193 pointer_inc           -> pointer_inc()
 *   lib/scanner.py:195-214 __get_bpp_mean        -> bpp_mean()
 *   lib/scanner.py:216-244 __filter_duplicates   -> orc_filter_duplicates()
 *   lib/bpprobs.py:6-18    BPProbs.add/prob_pair -> dense symmetric matrix, 0.0 when absent
 *   lib/gcx.py:4-13        GC.compute            -> orc_gc_compute()
 *   lib/evalues.py:12-36   select_gc / compute   -> orc_select_gc() / orc_evalue()
 *   lib/evalues.py:38-80   EValuesParser.parse   -> orc_evalues_load()
 *   lib/candidates.py:360-366 set_data_build     -> score = (pm - pr) / log(2.0)
 *   rmcalibrate.py:35-62,265-295 sample scoring  -> orc_calibrate()
 *
 * Pinned against the reference itself: tests/test_oracle_golden.py checks it
 * against the tests/golden/ json.gz fixtures, which tests/golden/make_golden.py produced
 * by running the reference's own Python code, and against the reference's
 * shipped result files (cluster_*.res, arich.out) and real dot.ps matrices.
 *
 * The model arrives as the text dump written by oracle/pyoracle.py
 * (`dump_model`), which mirrors the reference's Model attributes (nodes with
 * their per-key probability rows, ORDER, OFFSETS, pairs_list) rather than the
 * flattened device tables, so the two implementations share no table layout.
 */
#define _GNU_SOURCE
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXC 4          /* chains */
#define MAXL 64         /* positions per chain */
#define MAXKEY 8

/* ------------------------------------------------------------------ model */
typedef struct {
    char key[MAXKEY];
    int is_gc;
    int present[5];     /* A C G U . */
    double logp[5];
} prob_row;

typedef struct {
    int id, chain, ndx, nconds, conds[MAXKEY];
    int nrows;
    prob_row *rows;
} mnode;

typedef struct snode {
    int root, level, id, chain, ndx, nconds;
    const int *conds;
    const mnode *def;
    int oft;                    /* -1: None */
    struct snode *pred;
    struct snode **succs;
    int nsuccs;
    struct snode **parents;
    int has_min_dists, n_min_dists;
    int (*min_dists)[3];
    int cfg;                    /* OFFSETS index when the subtree holds one configuration */
    /* search state (reference lib/node.py:47-52) */
    char NT;
    int pos;                    /* -1: None */
    double prob_node, prob_rand, prob_best, rand_best;
    struct snode *succ_best;
} snode;

typedef struct {
    char name[64], version[32];
    int nchains, chain_len[MAXC], symmetric, sep_min;
    int nnodes;
    mnode *nodes;
    int *order;
    int ncfg;
    int *offsets;               /* [ncfg][MAXC][MAXL], -1 = None */
    int npaircfg;
    int *paircfg_n;             /* pairs per configuration */
    int (*pairs)[5];            /* c1 o1 c2 o2 ptype, concatenated */
    int *paircfg_start;
    snode *root;
} omodel;

static const char LETTERS[] = "ACGU.";

static int off_at(const omodel *m, int cfg, int chain, int ndx) {
    return m->offsets[(cfg * MAXC + chain) * MAXL + ndx];
}

static void resolve_parents(snode *self, snode *from, snode **parents) {
    /* reference lib/node.py:143-151: walk predecessors, the root excluded */
    for (snode *s = from; s && s->pred; s = s->pred)
        for (int k = 0; k < self->nconds; k++)
            if (s->id == self->conds[k]) parents[k] = s;
}

static snode *build_tree(omodel *m, const int *cfgs, int ncfgs, snode *pred, int level) {
    snode *s = calloc(1, sizeof(snode));
    s->root = level < 0;
    s->pred = pred;
    s->level = level;
    s->pos = -1;
    s->cfg = -1;
    if (!s->root) {
        const mnode *nd = &m->nodes[m->order[level]];
        s->def = nd;
        s->id = nd->id; s->chain = nd->chain; s->ndx = nd->ndx;
        s->nconds = nd->nconds; s->conds = nd->conds;
        s->oft = off_at(m, cfgs[0], s->chain, s->ndx);
        if (ncfgs == 1) {                       /* lib/node.py:70-79 */
            s->cfg = cfgs[0];
            s->has_min_dists = 1;
            s->min_dists = malloc(sizeof(int[3]) * m->nchains * m->nchains);
            for (int i = 0; i < m->nchains; i++) {
                int max_i = -1000000;
                for (int k = 0; k < m->chain_len[i]; k++) {
                    int o = off_at(m, cfgs[0], i, k);
                    if (o >= 0 && o > max_i) max_i = o;
                }
                for (int j = 0; j < m->nchains; j++)
                    if (i != j) {
                        int *d = s->min_dists[s->n_min_dists++];
                        d[0] = i; d[1] = j; d[2] = max_i + m->sep_min;
                    }
            }
        }
        if (s->nconds > 0) {                    /* lib/node.py:83-91 */
            s->parents = calloc(s->nconds, sizeof(snode *));
            resolve_parents(s, s, s->parents);
            for (int k = 0; k < s->nconds; k++)
                if (!s->parents[k]) {
                    fprintf(stderr, "oracle: unresolved parents in node %d\n", s->id);
                    abort();
                }
        }
    }
    int level_new = level + 1;
    if (level_new < m->nnodes) {                /* lib/node.py:97-117 */
        const mnode *nn = &m->nodes[m->order[level_new]];
        int *vals = malloc(sizeof(int) * ncfgs), nvals = 0;
        int *grp = malloc(sizeof(int) * ncfgs);
        for (int i = 0; i < ncfgs; i++) {
            int o = off_at(m, cfgs[i], nn->chain, nn->ndx), g = -1;
            for (int v = 0; v < nvals; v++) if (vals[v] == o) { g = v; break; }
            if (g < 0) { g = nvals; vals[nvals++] = o; }
            grp[i] = g;
        }
        s->succs = calloc(nvals, sizeof(snode *));
        for (int v = 0; v < nvals; v++) {
            int *sub = malloc(sizeof(int) * ncfgs), nsub = 0;
            for (int i = 0; i < ncfgs; i++) if (grp[i] == v) sub[nsub++] = cfgs[i];
            s->succs[s->nsuccs++] = build_tree(m, sub, nsub, s, level_new);
            free(sub);
        }
        free(vals); free(grp);
    }
    return s;
}

static void free_tree(snode *s) {
    if (!s) return;
    for (int i = 0; i < s->nsuccs; i++) free_tree(s->succs[i]);
    free(s->succs); free(s->parents); free(s->min_dists); free(s);
}

/* parse helpers for the text dump */
static char *next_line(char **cur) {
    char *s = *cur;
    if (!s || !*s) return NULL;
    char *e = strchr(s, '\n');
    if (e) { *e = 0; *cur = e + 1; } else { *cur = s + strlen(s); }
    return s;
}

void *orc_model_load(const char *dump) {
    char *buf = strdup(dump), *cur = buf, *line;
    omodel *m = calloc(1, sizeof(omodel));
    int node_i = -1, pc_i = -1, pair_i = 0;
    while ((line = next_line(&cur))) {
        char tag[32];
        int used = 0;
        if (sscanf(line, "%31s%n", tag, &used) != 1) continue;
        char *rest = line + used;
        if (!strcmp(tag, "MODEL")) sscanf(rest, "%63s %31s", m->name, m->version);
        else if (!strcmp(tag, "CHAINS")) {
            sscanf(rest, "%d%n", &m->nchains, &used); rest += used;
            for (int c = 0; c < m->nchains; c++) { sscanf(rest, "%d%n", &m->chain_len[c], &used); rest += used; }
        } else if (!strcmp(tag, "SYMMETRIC")) sscanf(rest, "%d", &m->symmetric);
        else if (!strcmp(tag, "SEPMIN")) sscanf(rest, "%d", &m->sep_min);
        else if (!strcmp(tag, "NODES")) {
            sscanf(rest, "%d", &m->nnodes);
            m->nodes = calloc(m->nnodes, sizeof(mnode));
            m->order = calloc(m->nnodes, sizeof(int));
        } else if (!strcmp(tag, "N")) {
            mnode *nd = &m->nodes[++node_i];
            sscanf(rest, "%d %d %d %d%n", &nd->id, &nd->chain, &nd->ndx, &nd->nconds, &used); rest += used;
            for (int k = 0; k < nd->nconds; k++) { sscanf(rest, "%d%n", &nd->conds[k], &used); rest += used; }
            sscanf(rest, "%d", &nd->nrows);
            nd->rows = calloc(nd->nrows ? nd->nrows : 1, sizeof(prob_row));
            nd->nrows = 0;
        } else if (!strcmp(tag, "R")) {
            mnode *nd = &m->nodes[node_i];
            prob_row *r = &nd->rows[nd->nrows++];
            char kind[8];
            sscanf(rest, "%7s %7s%n", r->key, kind, &used); rest += used;
            if (!strcmp(kind, "GC")) r->is_gc = 1;
            else for (int k = 0; k < 5; k++) {
                char tok[64];
                sscanf(rest, "%63s%n", tok, &used); rest += used;
                if (strcmp(tok, "-")) { r->present[k] = 1; r->logp[k] = strtod(tok, NULL); }
            }
        } else if (!strcmp(tag, "ORDER")) {
            for (int k = 0; k < m->nnodes; k++) { sscanf(rest, "%d%n", &m->order[k], &used); rest += used; }
        } else if (!strcmp(tag, "OFFSETS")) {
            sscanf(rest, "%d", &m->ncfg);
            m->offsets = malloc(sizeof(int) * m->ncfg * MAXC * MAXL);
            for (int i = 0; i < m->ncfg * MAXC * MAXL; i++) m->offsets[i] = -1;
        } else if (!strcmp(tag, "O")) {
            int cfg, chain;
            sscanf(rest, "%d %d%n", &cfg, &chain, &used); rest += used;
            for (int k = 0; k < m->chain_len[chain]; k++) {
                int v; sscanf(rest, "%d%n", &v, &used); rest += used;
                m->offsets[(cfg * MAXC + chain) * MAXL + k] = v;
            }
        } else if (!strcmp(tag, "PAIRCFG")) {
            int total;
            sscanf(rest, "%d %d", &m->npaircfg, &total);
            m->paircfg_n = calloc(m->npaircfg + 1, sizeof(int));
            m->paircfg_start = calloc(m->npaircfg + 1, sizeof(int));
            m->pairs = malloc(sizeof(int[5]) * (total ? total : 1));
        } else if (!strcmp(tag, "P")) {
            int n;
            sscanf(rest, "%d%n", &n, &used); rest += used;
            ++pc_i;
            m->paircfg_n[pc_i] = n;
            m->paircfg_start[pc_i] = pair_i;
            for (int k = 0; k < n; k++) {
                int *p = m->pairs[pair_i++];
                sscanf(rest, "%d %d %d %d %d%n", &p[0], &p[1], &p[2], &p[3], &p[4], &used); rest += used;
            }
        }
    }
    free(buf);
    int *all = malloc(sizeof(int) * m->ncfg);
    for (int i = 0; i < m->ncfg; i++) all[i] = i;
    m->root = build_tree(m, all, m->ncfg, NULL, -1);    /* lib/model.py:357-358 */
    free(all);
    return m;
}

void orc_model_free(void *h) {
    omodel *m = h;
    if (!m) return;
    free_tree(m->root);
    for (int i = 0; i < m->nnodes; i++) free(m->nodes[i].rows);
    free(m->nodes); free(m->order); free(m->offsets);
    free(m->paircfg_n); free(m->paircfg_start); free(m->pairs); free(m);
}

int orc_model_tree_size(void *h) {
    omodel *m = h;
    int n = 0;
    snode *stack[4096]; int sp = 0;
    stack[sp++] = m->root;
    while (sp) { snode *s = stack[--sp]; n++; for (int i = 0; i < s->nsuccs; i++) stack[sp++] = s->succs[i]; }
    return n - 1;
}

/* ------------------------------------------------------------------ GC */
/* gc256[c] = log frequency of byte c, has256[c] = key present in the dict */
typedef struct { double v[256]; unsigned char has[256]; } gcdict;

void orc_gc_compute(const char *seq, int64_t n, gcdict *out) {
    /* lib/gcx.py:4-13: float counts, log(count / len) */
    double cnt[256] = {0};
    memset(out, 0, sizeof(*out));
    for (int64_t i = 0; i < n; i++) cnt[(unsigned char)seq[i]] += 1.0;
    for (int c = 0; c < 256; c++)
        if (cnt[c] > 0) { out->has[c] = 1; out->v[c] = log(cnt[c] / (double)n); }
}

/* ------------------------------------------------------------------ eval */
typedef struct {
    const char *seq; int seqlen;
    const int *ptr;
    const gcdict *gc;
    double unknown_prob, cutoff;
    int *fault;                     /* set when the reference would raise (index/key error) */
} evalctx;

static int eval_node(snode *s, evalctx *x) {
    /* lib/node.py:176-277 */
    s->pos = -1; s->NT = 0; s->prob_node = 0.0; s->prob_rand = 0.0;
    if (s->has_min_dists)
        for (int k = 0; k < s->n_min_dists; k++) {
            int d = x->ptr[s->min_dists[k][1]] - x->ptr[s->min_dists[k][0]];
            if (d >= 0 && d < s->min_dists[k][2]) return 0;
        }
    if (!s->root) {
        if (s->oft < 0) { s->pos = -1; s->NT = '.'; }
        else {
            s->pos = x->ptr[s->chain] + s->oft;
            if (s->pos >= x->seqlen) { *x->fault = 1; return 0; }     /* IndexError in the reference */
            s->NT = x->seq[s->pos];
        }
        char key[MAXKEY + 1];
        if (s->nconds == 0) strcpy(key, "*");
        else { for (int k = 0; k < s->nconds; k++) key[k] = s->parents[k]->NT; key[s->nconds] = 0; }
        const prob_row *row = NULL;
        for (int r = 0; r < s->def->nrows; r++)
            if (!strcmp(s->def->rows[r].key, key)) { row = &s->def->rows[r]; break; }
        double prob_node;
        if (row && row->is_gc) {
            if (!x->gc->has[(unsigned char)s->NT]) { *x->fault = 1; return 0; }   /* KeyError */
            prob_node = x->gc->v[(unsigned char)s->NT];
        } else {
            prob_node = x->unknown_prob;
            if (row) {
                const char *lp = strchr(LETTERS, s->NT);
                if (lp && s->NT && row->present[lp - LETTERS]) prob_node = row->logp[lp - LETTERS];
            }
        }
        double prob_rand = x->gc->has[(unsigned char)s->NT] ? x->gc->v[(unsigned char)s->NT] : log(0.25);
        s->prob_node = s->pred->prob_node + prob_node;
        s->prob_rand = s->pred->prob_rand + prob_rand;
    }
    int ok = 0;
    if (s->prob_node >= (s->prob_rand + x->cutoff)) {
        s->succ_best = NULL; s->prob_best = -500; s->rand_best = -500;
        if (s->nsuccs == 0) {
            s->prob_best = s->prob_node; s->rand_best = s->prob_rand; ok = 1;
        } else {
            for (int i = 0; i < s->nsuccs; i++) {
                snode *c = s->succs[i];
                int cok = eval_node(c, x);
                if (cok && c->prob_best > s->prob_best) {
                    s->succ_best = c; s->prob_best = c->prob_best; s->rand_best = c->rand_best; ok = 1;
                }
            }
        }
    }
    return ok;
}

typedef struct {
    int ok, cfg, fault;
    double prob_model, prob_rand;
    char seqs[MAXC][MAXL];          /* letters with '.' for gaps, NUL terminated */
    int pos[MAXC][MAXL];            /* -1 = None */
} orc_solution;

static void get_solution(const omodel *m, orc_solution *out) {
    /* lib/node.py:153-174 */
    memset(out->seqs, 0, sizeof(out->seqs));
    for (int c = 0; c < MAXC; c++) for (int k = 0; k < MAXL; k++) out->pos[c][k] = 0;
    snode *s = m->root;
    out->cfg = -1;
    while (s) {
        if (!s->root) {
            out->seqs[s->chain][s->ndx] = s->NT;
            out->pos[s->chain][s->ndx] = s->pos;
            if (s->cfg >= 0) out->cfg = s->cfg;
        }
        s = s->succ_best;
    }
    out->prob_model = m->root->prob_best;
    out->prob_rand = m->root->rand_best;
}

int orc_eval(void *h, const char *seq, int seqlen, const int *ptr, const gcdict *gc,
             double unknown_prob, double cutoff, orc_solution *out) {
    omodel *m = h;
    int fault = 0;
    evalctx x = {seq, seqlen, ptr, gc, unknown_prob, cutoff, &fault};
    int ok = eval_node(m->root, &x);
    out->ok = ok && !fault; out->fault = fault;
    if (out->ok) get_solution(m, out);
    return out->ok;
}

/* ------------------------------------------------------------------ e-values */
typedef struct {
    int nclass, nrows;
    int nletters[512];
    char letters[512][8];
    double fracs[512][8];
    char (*keys)[16];
    double *vals;               /* [nrows][nclass] */
    int selected;               /* -1: None */
} oevalues;

void *orc_evalues_load(const char *path) {
    /* lib/evalues.py:38-80 */
    oevalues *e = calloc(1, sizeof(oevalues));
    e->selected = -1;
    FILE *f = fopen(path, "r");
    if (!f) return e;
    char *line = NULL; size_t cap = 0; ssize_t len;
    int state = 0, rowcap = 0;
    while ((len = getline(&line, &cap, f)) >= 0) {
        char *s = line;
        while (*s == ' ' || *s == '\t') s++;
        char *end = s + strlen(s);
        while (end > s && (end[-1] == '\n' || end[-1] == '\r' || end[-1] == ' ' || end[-1] == '\t')) *--end = 0;
        if (!*s || *s == '#') continue;
        if (strstr(s, "GC_CLASSES")) { state = 1; continue; }
        if (strstr(s, "E_VALUES")) { state = 2; continue; }
        if (state == 1) {
            int c = e->nclass++, used = 0;
            char name[64];
            sscanf(s, "%63s%n", name, &used); s += used;
            char l[8]; double v;
            while (sscanf(s, "%7s %lf%n", l, &v, &used) == 2) {
                /* a later duplicate letter overwrites (dict semantics) */
                int k, found = -1;
                for (k = 0; k < e->nletters[c]; k++) if (e->letters[c][k] == l[0]) found = k;
                if (found < 0) found = e->nletters[c]++;
                e->letters[c][found] = l[0]; e->fracs[c][found] = v;
                s += used;
            }
        } else if (state == 2) {
            if (e->nrows == rowcap) {
                rowcap = rowcap ? rowcap * 2 : 256;
                e->keys = realloc(e->keys, sizeof(char[16]) * rowcap);
                e->vals = realloc(e->vals, sizeof(double) * rowcap * e->nclass);
            }
            int used = 0, r = e->nrows;
            char key[16];
            sscanf(s, "%15s%n", key, &used); s += used;
            /* a repeated key overwrites the earlier row */
            for (int q = 0; q < e->nrows; q++) if (!strcmp(e->keys[q], key)) r = q;
            if (r == e->nrows) { strcpy(e->keys[r], key); e->nrows++; for (int c = 0; c < e->nclass; c++) e->vals[r * e->nclass + c] = 0.0; }
            for (int c = 0; c < e->nclass; c++) {
                char *endp; double v = strtod(s, &endp);
                if (endp == s) break;
                e->vals[r * e->nclass + c] = v; s = endp;
            }
        }
    }
    free(line); fclose(f);
    return e;
}

void orc_evalues_free(void *h) { oevalues *e = h; if (e) { free(e->keys); free(e->vals); free(e); } }
int orc_evalues_nclass(void *h) { return ((oevalues *)h)->nclass; }
int orc_evalues_nrows(void *h) { return ((oevalues *)h)->nrows; }

int orc_select_gc(void *h, const gcdict *gc) {
    /* lib/evalues.py:12-23 */
    oevalues *e = h;
    double min_d = 0; int have = 0;
    for (int i = 0; i < e->nclass; i++) {
        double x = 0.0;
        for (int k = 0; k < e->nletters[i]; k++) {
            unsigned char c = (unsigned char)e->letters[i][k];
            double g = gc->has[c] ? gc->v[c] : 0.0;
            x += pow(e->fracs[i][k] - pow(M_E, g), 2.0);
        }
        double d = sqrt(x);
        if (!have || d < min_d) { min_d = d; e->selected = i; have = 1; }
    }
    return e->selected;
}

void orc_set_selected(void *h, int cls) { ((oevalues *)h)->selected = cls; }

double orc_evalue(void *h, double score) {
    /* lib/evalues.py:25-36.  "%.1f" % round(score, 1): both steps are correctly rounded
       decimal conversions of the binary value (ties to even), so one glibc "%.1f" of the
       score itself yields the same string. */
    oevalues *e = h;
    if (e->selected < 0) return 1.0;
    char key[64];
    snprintf(key, sizeof key, "%.1f", score);
    if (!strcmp(key, "-0.0")) strcpy(key, "0.0");
    for (int r = 0; r < e->nrows; r++)
        if (!strcmp(e->keys[r], key)) return e->vals[r * e->nclass + e->selected];
    return 0.0;
}

/* ------------------------------------------------------------------ windows */
int64_t orc_windows(int64_t seqlen, int win_len, int win_step, int64_t *starts, int64_t cap) {
    /* lib/scanner.py:80-111: returns the number of windows; fills starts[] up to cap.
       Window k covers [starts[k], min(starts[k]+win_len, seqlen)). */
    int64_t n = 0, ptr = 0;
    int ok = 1;
    while (ok) {
        int64_t end = ptr + win_len;
        if (end >= seqlen) { ptr = seqlen - win_len > 0 ? seqlen - win_len : 0; ok = 0; }
        if (n < cap) starts[n] = ptr;
        n++;
        ptr += win_step;
    }
    return n;
}

/* ------------------------------------------------------------------ search */
typedef struct {
    int32_t seq, model;
    int64_t coord0, coord1;         /* window coordinates after update_chains_pos */
    int32_t p[MAXC];                /* window-relative pointers */
    int32_t cfg;
    double prob_model, prob_rand, score, evalue;
    char seqs[MAXC][MAXL];
    int64_t pos[MAXC][MAXL];        /* absolute, -1 = None */
} orc_hit;

typedef struct { orc_hit *v; int64_t n, cap; } hitvec;

static void hv_push(hitvec *hv, const orc_hit *h) {
    if (hv->n == hv->cap) { hv->cap = hv->cap ? hv->cap * 2 : 64; hv->v = realloc(hv->v, hv->cap * sizeof(orc_hit)); }
    hv->v[hv->n++] = *h;
}

static int pointer_inc(const omodel *m, int seq_len, int *ptr) {
    /* lib/scanner.py:161-193 */
    int i = m->nchains - 1;
    for (;;) {
        ptr[i] += 1;
        if (ptr[i] + m->chain_len[i] >= seq_len) {
            i -= 1;
            if (i < 0) return 0;
            if (m->symmetric) ptr[i + 1] = ptr[i] + 1; else ptr[i + 1] = 0;
        } else break;
    }
    return 1;
}

static double bpp_mean(const omodel *m, const double *mat, int w, const int *ptr) {
    /* lib/scanner.py:195-214; mat is the dense symmetric matrix, absent = 0.0 */
    double sum_prob = 0.0; int sum_pairs = 0;
    for (int c = 0; c < m->npaircfg; c++) {
        int s = m->paircfg_start[c];
        for (int k = 0; k < m->paircfg_n[c]; k++) {
            const int *p = m->pairs[s + k];
            if (p[4] == 1) {
                int b1 = ptr[p[0]] + p[1], b2 = ptr[p[2]] + p[3];
                double prob = (b1 < w && b2 < w && b1 >= 0 && b2 >= 0) ? mat[(size_t)b1 * w + b2] : 0.0;
                if (prob == 0) break;
                sum_prob += prob; sum_pairs += 1;
            }
        }
    }
    if (sum_pairs == 0) return 0.0;
    return sum_prob / (double)sum_pairs;
}

/* One window against every model (lib/scanner.py:118-159).  models/evalues arrays are parallel. */
static void search_window(void **models, void **evalues, int nmodels, int seq_id,
                          const char *win, int wlen, int64_t ptr0, const gcdict *gc,
                          const uint16_t *bi, const uint16_t *bj, const double *bp, int64_t nnz,
                          double min_bpp_mean, double unknown_prob, double cutoff,
                          hitvec *hits, int64_t tries[2], int *fault, double *mat) {
    if (wlen <= 0) return;
    memset(mat, 0, sizeof(double) * (size_t)wlen * wlen);
    for (int64_t k = 0; k < nnz; k++)
        if (bp[k] != 0.0 && bi[k] < wlen && bj[k] < wlen) {
            mat[(size_t)bi[k] * wlen + bj[k]] = bp[k];
            mat[(size_t)bj[k] * wlen + bi[k]] = bp[k];
        }
    for (int mi = 0; mi < nmodels; mi++) {
        omodel *m = models[mi];
        if (evalues && evalues[mi]) orc_select_gc(evalues[mi], gc);
        int ptr[MAXC] = {0, 0, 0, 0};
        for (;;) {
            tries[0] += 1;
            if (bpp_mean(m, mat, wlen, ptr) >= min_bpp_mean) {
                tries[1] += 1;
                orc_solution sol;
                int ok = orc_eval(m, win, wlen, ptr, gc, unknown_prob, cutoff, &sol);
                if (sol.fault) *fault = 1;
                if (ok) {
                    orc_hit h;
                    memset(&h, 0, sizeof h);
                    h.seq = seq_id; h.model = mi;
                    h.coord0 = ptr0; h.coord1 = ptr0 + wlen;
                    for (int c = 0; c < m->nchains; c++) h.p[c] = ptr[c];
                    h.cfg = sol.cfg;
                    h.prob_model = sol.prob_model; h.prob_rand = sol.prob_rand;
                    h.score = (sol.prob_model - sol.prob_rand) / log(2.0);      /* lib/candidates.py:365 */
                    h.evalue = (evalues && evalues[mi]) ? orc_evalue(evalues[mi], h.score) : 1.0;
                    memcpy(h.seqs, sol.seqs, sizeof h.seqs);
                    for (int c = 0; c < MAXC; c++)
                        for (int k = 0; k < MAXL; k++)
                            h.pos[c][k] = sol.pos[c][k] < 0 ? -1 : sol.pos[c][k] + ptr0;  /* update_chains_pos */
                    hv_push(hits, &h);
                }
            }
            if (!pointer_inc(m, wlen, ptr)) break;
        }
    }
}

int64_t orc_filter_duplicates(orc_hit *c, int64_t n, int nchains) {
    /* lib/scanner.py:216-244, literally; returns the new length (order preserved) */
    char *dead = calloc(n ? n : 1, 1);
    for (int64_t i = 0; i < n; i++) {
        if (dead[i]) continue;
        for (int64_t j = i + 1; j < n; j++) {
            if (dead[j]) continue;
            int dup = 1;
            if (c[i].coord0 < c[j].coord1 && c[j].coord0 < c[i].coord1 &&
                c[i].seq == c[j].seq && c[i].model == c[j].model) {
                for (int k = 0; k < nchains; k++)
                    if (c[i].pos[k][0] != c[j].pos[k][0]) { dup = 0; break; }
            } else dup = 0;
            if (dup) {
                if (c[i].score > c[j].score) {
                    /* the reference would now compare bpp (None) with 0 and raise */
                    fprintf(stderr, "oracle: duplicate with a larger earlier score — undefined in the reference\n");
                    abort();
                }
                dead[i] = 1;
                break;
            }
        }
    }
    int64_t w = 0;
    for (int64_t i = 0; i < n; i++) if (!dead[i]) c[w++] = c[i];
    free(dead);
    return w;
}

/*
 * Scan one sequence (lib/scanner.py:29-68 for one element of the list): slide or whole-sequence
 * search, then duplicate removal.  bpp for window k is the coordinate list
 * [bpp_off[k], bpp_off[k+1]) of (bi, bj, bp).  Returns a malloc'd hit array in *out.
 */
int64_t orc_scan_sequence(void **models, void **evalues, int nmodels, int seq_id,
                          const char *seq, int64_t seqlen, int win_len, int win_step,
                          const int64_t *bpp_off, const uint16_t *bi, const uint16_t *bj, const double *bp,
                          double min_bpp_mean, double unknown_prob, double cutoff,
                          orc_hit **out, int64_t tries[2], int64_t *tries_per_window, int *fault,
                          const gcdict *gc_given /* Scanner.scan(seqs, gc): lib/scanner.py:29,74-75 */) {
    hitvec hv = {0};
    gcdict gc;
    tries[0] = tries[1] = 0; *fault = 0;
    int nch = nmodels ? ((omodel *)models[0])->nchains : 2;
    if (gc_given) gc = *gc_given;
    else if (seqlen > 0) orc_gc_compute(seq, seqlen, &gc); else memset(&gc, 0, sizeof gc);
    if (win_len == 0 || win_step == 0) {
        if (seqlen > 0) {
            double *mat = malloc(sizeof(double) * (size_t)seqlen * seqlen);
            search_window(models, evalues, nmodels, seq_id, seq, (int)seqlen, 0, &gc,
                          bi + bpp_off[0], bj + bpp_off[0], bp + bpp_off[0], bpp_off[1] - bpp_off[0],
                          min_bpp_mean, unknown_prob, cutoff, &hv, tries, fault, mat);
            free(mat);
        }
        if (tries_per_window) { tries_per_window[0] = tries[0]; tries_per_window[1] = tries[1]; }
    } else {
        int64_t nwin = orc_windows(seqlen, win_len, win_step, NULL, 0);
        int64_t *starts = malloc(sizeof(int64_t) * nwin);
        orc_windows(seqlen, win_len, win_step, starts, nwin);
        double *mat = malloc(sizeof(double) * (size_t)win_len * win_len);
        for (int64_t k = 0; k < nwin; k++) {
            int64_t s = starts[k];
            int wl = (int)((s + win_len <= seqlen ? s + win_len : seqlen) - s);
            int64_t t[2] = {0, 0};
            search_window(models, evalues, nmodels, seq_id, seq + s, wl, s, &gc,
                          bi + bpp_off[k], bj + bpp_off[k], bp + bpp_off[k], bpp_off[k + 1] - bpp_off[k],
                          min_bpp_mean, unknown_prob, cutoff, &hv, t, fault, mat);
            tries[0] += t[0]; tries[1] += t[1];
            if (tries_per_window) { tries_per_window[2 * k] = t[0]; tries_per_window[2 * k + 1] = t[1]; }
        }
        free(mat); free(starts);
    }
    hv.n = orc_filter_duplicates(hv.v, hv.n, nch);
    *out = hv.v;
    return hv.n;
}

void orc_free(void *p) { free(p); }
int orc_sizeof_hit(void) { return (int)sizeof(orc_hit); }
int orc_sizeof_solution(void) { return (int)sizeof(orc_solution); }
int orc_sizeof_gcdict(void) { return (int)sizeof(gcdict); }

/* ------------------------------------------------------------------ synthetic bpp (C twin of
 * rmdetect_b200/synth.py; used by the CPU-baseline leg so that it does not sit in numpy) */
static uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static int hash_byte(char c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'U': return 3; default: return (unsigned char)c; }
}
static int can_pair(int a, int b) {
    if (a > 3 || b > 3) return 0;
    int s = a + b;
    return (s == 3 && a != b) || (s == 5 && a != b && (a < b ? a : b) == 2);
}
static uint64_t cell_hash(uint64_t key, uint64_t i, uint64_t j, uint64_t salt) {
    uint64_t v = (i << 34) | (j << 4) | salt;
    return mix64(key + 0x9E3779B97F4A7C15ULL * (v + 1));
}
uint64_t orc_window_key(const char *win, int n, uint64_t seed) {
    uint64_t key = seed;
    for (int b = 0; b < n; b += 8) {
        uint64_t w = 0;
        for (int k = 0; k < 8 && b + k < n; k++) w |= (uint64_t)hash_byte(win[b + k]) << (8 * k);
        key = mix64(key ^ w);
    }
    return mix64(key ^ (uint64_t)n);
}
int64_t orc_synth_bpp(const char *win, int n, uint64_t seed, uint16_t *oi, uint16_t *oj, double *op, int64_t cap) {
    static const int len_tab[16] = {1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 6, 7, 8};
    static const int64_t band_lo[3] = {3162278, 31622777, 316227767};
    static const int64_t band_hi[3] = {31622777, 316227767, 1000000001};
    if (n < 5) return 0;
    uint64_t key = orc_window_key(win, n, seed);
    int64_t cnt = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 4; j < n; j++) {
            int64_t best = 0;
            for (int t = 0; t < 8; t++) {
                int si = i - t, sj = j + t;
                if (si < 0 || sj >= n) break;
                if (!can_pair(hash_byte(win[si]), hash_byte(win[sj]))) break;
                uint64_t h0 = cell_hash(key, si, sj, 0);
                if ((h0 >> 40) >= (31u << 16)) continue;
                if (len_tab[(h0 >> 8) & 15] <= t) continue;
                int64_t bu = (h0 >> 16) & 0xFFFF;
                int band = bu < 49152 ? 0 : (bu < 60000 ? 1 : 2);
                uint64_t h1 = cell_hash(key, si, sj, 1);
                int64_t base = band_lo[band] + (int64_t)(h1 % (uint64_t)(band_hi[band] - band_lo[band]));
                uint64_t h2 = cell_hash(key, i, j, 2 + t);
                int64_t half = base >> 1;
                int64_t val = half + ((half * (int64_t)(h2 & 1023)) >> 10);
                if (val < 3162278) val = 3162278;
                if (val > best) best = val;
            }
            if (best > 0) {
                if (cnt < cap) { double r = (double)best / 1e9; oi[cnt] = i; oj[cnt] = j; op[cnt] = r * r; }
                cnt++;
            }
        }
    return cnt;
}

/*
 * CPU-baseline driver: scan windows [w0, w1) of one ACGU sequence with the synthetic bpp twin,
 * OpenMP over windows (the reference itself is single-threaded; this is "all the host threads it
 * can use").  GC is taken over the whole sequence, as slide does.  Only counts are returned.
 */
int64_t orc_bench_scan(int nmodels, const char **model_dumps,
                       const char *seq, int64_t seqlen, int win_len, int win_step,
                       int64_t w0, int64_t w1, uint64_t seed, int nthreads,
                       double min_bpp_mean, double unknown_prob, double cutoff, int64_t tries[2]) {
    gcdict gc;
    orc_gc_compute(seq, seqlen, &gc);
    int64_t nwin = orc_windows(seqlen, win_len, win_step, NULL, 0);
    int64_t *starts = malloc(sizeof(int64_t) * nwin);
    orc_windows(seqlen, win_len, win_step, starts, nwin);
    if (w1 > nwin) w1 = nwin;
    int64_t t0 = 0, t1 = 0, nh = 0;
#pragma omp parallel num_threads(nthreads) reduction(+ : t0, t1, nh)
    {
        /* the search tree holds evaluation state, so every thread builds its own models */
        void **ms = malloc(sizeof(void *) * nmodels);
        for (int i = 0; i < nmodels; i++) ms[i] = orc_model_load(model_dumps[i]);
        double *mat = malloc(sizeof(double) * (size_t)win_len * win_len);
        int64_t cap = (int64_t)win_len * win_len / 2 + 16;
        uint16_t *bi = malloc(sizeof(uint16_t) * cap), *bj = malloc(sizeof(uint16_t) * cap);
        double *bp = malloc(sizeof(double) * cap);
#pragma omp for schedule(dynamic, 4)
        for (int64_t k = w0; k < w1; k++) {
            int64_t s = starts[k];
            int wl = (int)((s + win_len <= seqlen ? s + win_len : seqlen) - s);
            int64_t nnz = orc_synth_bpp(seq + s, wl, seed, bi, bj, bp, cap);
            hitvec hv = {0};
            int64_t t[2] = {0, 0};
            int fault = 0;
            search_window(ms, NULL, nmodels, 0, seq + s, wl, s, &gc, bi, bj, bp, nnz,
                          min_bpp_mean, unknown_prob, cutoff, &hv, t, &fault, mat);
            t0 += t[0]; t1 += t[1]; nh += hv.n;
            free(hv.v);
        }
        free(mat); free(bi); free(bj); free(bp);
        for (int i = 0; i < nmodels; i++) orc_model_free(ms[i]);
        free(ms);
    }
    free(starts);
    tries[0] = t0; tries[1] = t1;
    return nh;
}


/*
 * The same driver with the two arms of bench.py made like for like: the stand-in bpp lists of a block of windows
 * are generated first (their time goes to secs[0], as ViennaRNA's would), the scan of the block is timed apart
 * (secs[1]).  The GC table is the caller's (letter counts of the WHOLE chromosome, lib/scanner.py:74-75), so a
 * prefix sample scores exactly as the full scan does.  checksum: order-independent sum over the hits before
 * duplicate removal — the formula of rmdetect_b200/genome.py hit_checksum (absolute chain positions, model, gap
 * configuration, score bits).
 */
int64_t orc_bench_scan2(int nmodels, const char **model_dumps,
                        const char *seq, int64_t seqlen, int win_len, int win_step,
                        int64_t w0, int64_t w1, uint64_t seed, int nthreads, const int64_t counts[4],
                        double min_bpp_mean, double unknown_prob, double cutoff,
                        int64_t tries[2], uint64_t *checksum, double secs[2]) {
    gcdict gc;
    if (counts) {                                  /* GC.compute on the whole sequence the sample was cut from */
        memset(&gc, 0, sizeof gc);
        double total = (double)(counts[0] + counts[1] + counts[2] + counts[3]);
        for (int k = 0; k < 4; k++)
            if (counts[k]) { gc.has[(unsigned char)LETTERS[k]] = 1; gc.v[(unsigned char)LETTERS[k]] = log((double)counts[k] / total); }
    } else orc_gc_compute(seq, seqlen, &gc);
    int64_t nwin = orc_windows(seqlen, win_len, win_step, NULL, 0);
    int64_t *starts = malloc(sizeof(int64_t) * nwin);
    orc_windows(seqlen, win_len, win_step, starts, nwin);
    if (w1 > nwin) w1 = nwin;
    const int64_t BLOCK = 8192;
    const int64_t cap = (int64_t)win_len * win_len / 2 + 16;
    int64_t *boff = malloc(sizeof(int64_t) * (BLOCK + 1));
    uint16_t *bi = NULL, *bj = NULL; double *bp = NULL; int64_t bcap = 0;
    int64_t *cnt = malloc(sizeof(int64_t) * BLOCK);
    int64_t t0 = 0, t1 = 0, nh = 0; uint64_t sum = 0;
    secs[0] = secs[1] = 0.0;
    /* every thread builds its own models: the search tree holds evaluation state */
    void ***ms = malloc(sizeof(void **) * nthreads);
    double **mats = malloc(sizeof(double *) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        ms[t] = malloc(sizeof(void *) * nmodels);
        for (int i = 0; i < nmodels; i++) ms[t][i] = orc_model_load(model_dumps[i]);
        mats[t] = malloc(sizeof(double) * (size_t)win_len * win_len);
    }
    for (int64_t b0 = w0; b0 < w1; b0 += BLOCK) {
        const int64_t b1 = b0 + BLOCK < w1 ? b0 + BLOCK : w1, nb = b1 - b0;
        double ta = omp_get_wtime();
        /* pass 1: sizes; pass 2: lists (the generator is deterministic) */
#pragma omp parallel num_threads(nthreads)
        {
            uint16_t *ti = malloc(sizeof(uint16_t) * cap), *tj = malloc(sizeof(uint16_t) * cap);
            double *tp = malloc(sizeof(double) * cap);
#pragma omp for schedule(dynamic, 16)
            for (int64_t k = 0; k < nb; k++) {
                int64_t s = starts[b0 + k];
                int wl = (int)((s + win_len <= seqlen ? s + win_len : seqlen) - s);
                cnt[k] = orc_synth_bpp(seq + s, wl, seed, ti, tj, tp, cap);
            }
            free(ti); free(tj); free(tp);
        }
        boff[0] = 0;
        for (int64_t k = 0; k < nb; k++) boff[k + 1] = boff[k] + cnt[k];
        if (boff[nb] > bcap) {
            bcap = boff[nb] + boff[nb] / 8 + 1024;
            free(bi); free(bj); free(bp);
            bi = malloc(sizeof(uint16_t) * bcap); bj = malloc(sizeof(uint16_t) * bcap); bp = malloc(sizeof(double) * bcap);
        }
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
        for (int64_t k = 0; k < nb; k++) {
            int64_t s = starts[b0 + k];
            int wl = (int)((s + win_len <= seqlen ? s + win_len : seqlen) - s);
            orc_synth_bpp(seq + s, wl, seed, bi + boff[k], bj + boff[k], bp + boff[k], cnt[k]);
        }
        double tb = omp_get_wtime();
#pragma omp parallel num_threads(nthreads) reduction(+ : t0, t1, nh, sum)
        {
            const int me = omp_get_thread_num();
#pragma omp for schedule(dynamic, 4)
            for (int64_t k = 0; k < nb; k++) {
                int64_t s = starts[b0 + k];
                int wl = (int)((s + win_len <= seqlen ? s + win_len : seqlen) - s);
                hitvec hv = {0};
                int64_t t[2] = {0, 0};
                int fault = 0;
                search_window(ms[me], NULL, nmodels, 0, seq + s, wl, s, &gc, bi + boff[k], bj + boff[k], bp + boff[k], cnt[k],
                              min_bpp_mean, unknown_prob, cutoff, &hv, t, &fault, mats[me]);
                t0 += t[0]; t1 += t[1]; nh += hv.n;
                for (int64_t q = 0; q < hv.n; q++) {
                    const orc_hit *h = &hv.v[q];
                    uint64_t key = (uint64_t)(h->coord0 + h->p[0]) * 0x9E3779B97F4A7C15ULL;
                    key ^= (uint64_t)(h->coord0 + h->p[1]) * 0xC2B2AE3D27D4EB4FULL;
                    key ^= ((uint64_t)h->model << 40) ^ ((uint64_t)h->cfg << 48);
                    key = (key ^ (key >> 29)) * 0xBF58476D1CE4E5B9ULL;
                    uint64_t bits;
                    memcpy(&bits, &h->score, sizeof bits);
                    sum += key + bits;
                }
                free(hv.v);
            }
        }
        double tc = omp_get_wtime();
        secs[0] += tb - ta; secs[1] += tc - tb;
    }
    for (int t = 0; t < nthreads; t++) {
        for (int i = 0; i < nmodels; i++) orc_model_free(ms[t][i]);
        free(ms[t]); free(mats[t]);
    }
    free(ms); free(mats); free(starts); free(boff); free(cnt); free(bi); free(bj); free(bp);
    tries[0] = t0; tries[1] = t1;
    if (checksum) *checksum = sum;
    return nh;
}

/* ------------------------------------------------------------------ calibration */
/*
 * rmcalibrate.py:265-295 for a batch of samples.  samples: n strings of L letters (row-major,
 * no terminators).  classes: nclass x 4 fractions in the order A C G U.  hist: nclass x nbins
 * doubles, bin k = key round(score, 1) == k/10; accumulated in sample order like the dicts.
 * Returns the number of samples whose eval succeeded.
 */
int64_t orc_calibrate(void *h, const char *samples, int64_t n, int L, const double *classes, int nclass,
                      double unknown_prob, double cutoff, double *hist, int nbins, int *overflow) {
    omodel *m = h;
    static const char SPACER[] = "CCCCCC";
    int L0 = m->chain_len[0];
    int ptr[MAXC] = {0, L0 + 6, 0, 0};
    gcdict neutral;
    memset(&neutral, 0, sizeof neutral);
    for (const char *c = "ACGU"; *c; c++) { neutral.has[(unsigned char)*c] = 1; neutral.v[(unsigned char)*c] = log(0.25); }
    double K = L * log(4.0);
    char seq[2 * MAXL + 16];
    int64_t nok = 0;
    *overflow = 0;
    for (int64_t s = 0; s < n; s++) {
        const char *smp = samples + s * L;
        memcpy(seq, smp, L0); memcpy(seq + L0, SPACER, 6); memcpy(seq + L0 + 6, smp + L0, L - L0);
        int slen = L + 6;
        orc_solution sol;
        int ok = orc_eval(m, seq, slen, ptr, &neutral, unknown_prob, cutoff, &sol);
        char used[2 * MAXL]; int nused = 0;
        if (ok) {
            nok++;
            for (int c = 0; c < m->nchains; c++)
                for (int k = 0; k < m->chain_len[c]; k++)
                    if (sol.seqs[c][k] != '.') used[nused++] = sol.seqs[c][k];
        }
        for (int i = 0; i < nclass; i++) {
            const double *gc = classes + 4 * i;
            double score = 0.0;
            if (ok) {                                       /* get_gc_scores :47-62 */
                double p = 0.0;
                for (int k = 0; k < nused; k++) p += log(gc[strchr(LETTERS, used[k]) - LETTERS]);
                score = (sol.prob_model - p) / log(2.0);
                if (!(score > 0.0)) score = 0.0;            /* max(x, 0.0) */
            }
            double p = 0.0;                                 /* get_gc_weights :35-45 */
            for (int k = 0; k < L; k++) p += log(gc[strchr(LETTERS, smp[k]) - LETTERS]);
            double weight = pow(M_E, p + K);
            char key[64];
            snprintf(key, sizeof key, "%.1f", score);       /* round(score, 1) as a tenths index */
            long bin = lround(strtod(key, NULL) * 10.0);
            if (bin < 0 || bin >= nbins) { *overflow = 1; continue; }
            hist[(size_t)i * nbins + bin] += weight;
        }
    }
    return nok;
}

/* Test support for the device's expansion of the compact bpp transport (rmdetect_b200/csrc/device_common.cuh:
   transport_root): q * 1e-9 corrected by its exact residual equals the IEEE quotient q / 1e9 for EVERY q below 2^30.
   The three operations are correctly rounded IEEE operations on the host and on the GPU alike, so the exhaustive check
   here settles it for both.  Returns the number of mismatches (0). */
__attribute__((target("fma"))) int64_t orc_check_div1e9(void) {
    int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t q = 0; q < ((int64_t)1 << 30); q++) {
        const double x = (double)q, want = x / 1e9;
        const double y = x * 1e-9;
        const double got = fma(fma(-y, 1e9, x), 1e-9, y);
        if (got != want) bad++;
    }
    return bad;
}
