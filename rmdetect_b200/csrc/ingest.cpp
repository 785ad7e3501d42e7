// ingest.cpp — host-side ingest of RNAfold output (no device work).
//
//   rmd_parse_dot_ps   replaces lib/rnatools.py:176-205 (parse_bpprobs) of the reference: `i j sqrt(p) ubox`
//                      lines of a dot.ps become the sorted upper-triangle coordinate list the kernels consume;
//                      p = sqrt_p ** 2.0 through libm's pow, as CPython computes it for the reference.
//   rmd_compact_bpp    6-byte transport form of an entry (see include/rmdetect_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/rmdetect_b200.h"

namespace {
inline uint64_t bits_of(double x) { uint64_t b; memcpy(&b, &x, 8); return b; }

// q | (c + 1) << 30 with p == nextafter^c((q / 1e9)^2); false when p has no such form
inline bool compact_value(double p, uint32_t *out) {
    if (!(p > 0.0) || !(p <= 1.0000001)) return false;
    const double r = std::nearbyint(std::sqrt(p) * 1e9);
    if (!(r >= 1.0) || r >= 1073741824.0) return false;
    const double x = r / 1e9, sq = x * x;
    const int64_t d = (int64_t)bits_of(p) - (int64_t)bits_of(sq);      // positive doubles: ulp distance
    if (d < -1 || d > 1) return false;
    *out = (uint32_t)r | ((uint32_t)(d + 1) << 30);
    return true;
}
}  // namespace

extern "C" int rmd_compact_bpp(const uint16_t *i, const uint16_t *j, const double *p, int64_t n,
                               uint16_t *ij, uint32_t *q, int64_t *first_bad) {
    if (n < 0 || (n > 0 && (!i || !j || !p || !ij || !q))) return RMD_E_INVALID;
    for (int64_t k = 0; k < n; k++) {
        uint32_t v;
        if (i[k] > 255 || j[k] > 255 || !compact_value(p[k], &v)) {
            if (first_bad) *first_bad = k;
            return RMD_E_INVALID;
        }
        ij[k] = (uint16_t)(i[k] | (j[k] << 8));
        q[k] = v;
    }
    return RMD_OK;
}

extern "C" int rmd_parse_dot_ps(const char *text, int64_t len, uint16_t *oi, uint16_t *oj, double *op, uint32_t *oq,
                                int64_t cap, int64_t *n_out, char *seq, int64_t seq_cap) {
    if (!text || len < 0 || !n_out) return RMD_E_INVALID;
    struct Entry { uint32_t key; uint32_t order; double p; };
    std::vector<Entry> entries;
    std::vector<char> letters;
    bool in_seq = false;
    const char *cur = text, *end = text + len;
    volatile double two = 2.0;                                  // keep the libm call: pow(x, 2.0) is not always x * x
    while (cur < end) {
        const char *eol = (const char *)memchr(cur, '\n', (size_t)(end - cur));
        if (!eol) eol = end;
        const char *a = cur, *b = eol;
        while (a < b && (*a == ' ' || *a == '\t' || *a == '\r')) a++;
        while (b > a && (b[-1] == ' ' || b[-1] == '\t' || b[-1] == '\r')) b--;
        const size_t ll = (size_t)(b - a);
        if (ll >= 9 && memcmp(a, "/sequence", 9) == 0) { in_seq = true; letters.clear(); }
        else if (in_seq) {
            if (ll > 0 && *a == ')') in_seq = false;
            else { const char *e = b; while (e > a && e[-1] == '\\') e--; letters.insert(letters.end(), a, e); }
        } else if (ll > 4 && memcmp(b - 4, "ubox", 4) == 0 && *a != '%' && *a != '/') {
            // exactly four whitespace-separated tokens: i j sqrt(p) ubox
            char buf[128];
            if (ll < sizeof buf) {
                memcpy(buf, a, ll); buf[ll] = 0;
                char *tok[5]; int nt = 0;
                for (char *t = strtok(buf, " \t"); t && nt < 5; t = strtok(nullptr, " \t")) tok[nt++] = t;
                if (nt == 4) {
                    char *e1, *e2, *e3;
                    const long i1 = strtol(tok[0], &e1, 10), j1 = strtol(tok[1], &e2, 10);
                    const double sq = strtod(tok[2], &e3);
                    if (!*e1 && !*e2 && !*e3 && i1 >= 1 && j1 >= 1 && i1 <= 65535 && j1 <= 65535 && i1 != j1) {
                        const double p = pow(sq, two);
                        if (p != 0.0) {
                            const uint32_t lo = (uint32_t)std::min(i1, j1) - 1, hi = (uint32_t)std::max(i1, j1) - 1;
                            entries.push_back(Entry{(lo << 16) | hi, (uint32_t)entries.size(), p});
                        }
                    }
                }
            }
        }
        cur = eol + 1;
    }
    // a later line for the same pair overwrites an earlier one (BPProbs.add, lib/bpprobs.py:6-15)
    std::sort(entries.begin(), entries.end(), [](const Entry &x, const Entry &y) { return x.key != y.key ? x.key < y.key : x.order < y.order; });
    size_t m = 0;
    for (size_t k = 0; k < entries.size(); k++)
        if (k + 1 == entries.size() || entries[k + 1].key != entries[k].key) entries[m++] = entries[k];
    entries.resize(m);
    *n_out = (int64_t)m;
    if (seq && seq_cap > 0) {
        const size_t take = std::min<size_t>(letters.size(), (size_t)seq_cap - 1);
        memcpy(seq, letters.data(), take);
        seq[take] = 0;
    }
    if ((int64_t)m > cap) return RMD_E_LIMIT;
    for (size_t k = 0; k < m; k++) {
        if (oi) oi[k] = (uint16_t)(entries[k].key >> 16);
        if (oj) oj[k] = (uint16_t)(entries[k].key & 0xffff);
        if (op) op[k] = entries[k].p;
        if (oq && !compact_value(entries[k].p, &oq[k])) oq[k] = 0;      // 0: no compact form (callers fall back to fp64)
    }
    return RMD_OK;
}
