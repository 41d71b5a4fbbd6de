/*
 * fzoracle.c -- CPU restatement of the fuzzysearch hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA library in fuzzysearch_b200/csrc.  It is a
 * plain-C restatement of the reference's *pure-Python* algorithms (config "P" in SURVEY.md
 * section 8c), written so that every function follows the cited reference lines statement by
 * statement (including their quirks).  Nothing under fuzzysearch_b200/ may link, import or
 * call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do.
 *
 * Pinning: tests/test_oracle_golden.py checks every function here against fixtures generated
 * by importing the real reference (tests/golden/gen_golden.py, run in the build container where
 * /root/reference exists) -- the reference's own test tables plus seeded fuzz.
 *
 * All citations are relative to /root/reference/src/fuzzysearch/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FZO_API __attribute__((visibility("default")))

typedef struct {
    int64_t start, end, dist;
} fzo_match;

typedef struct {
    fzo_match *v;
    int64_t n, cap;
} fzo_list;

static int push(fzo_list *L, int64_t s, int64_t e, int64_t d) {
    if (L->n == L->cap) {
        int64_t ncap = L->cap ? L->cap * 2 : 64;
        fzo_match *nv = (fzo_match *)realloc(L->v, (size_t)ncap * sizeof(fzo_match));
        if (!nv) return -1;
        L->v = nv;
        L->cap = ncap;
    }
    L->v[L->n].start = s;
    L->v[L->n].end = e;
    L->v[L->n].dist = d;
    L->n++;
    return 0;
}

FZO_API void fzo_free(void *p) { free(p); }

static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * search_exact (search_exact.py:22-56, bytes path: sequence.find(sub, start, end) loop).
 * clamp() of the two indexes as in search_exact.py:29-30.  Calls cb for each index ascending.
 * ---------------------------------------------------------------------------------------- */
static void clamp_window(int64_t n, int64_t *start, int64_t *end) {
    int64_t s = imax(0, imin(*start, n));  /* clamp(start, 0, n)   search_exact.py:29 */
    int64_t e = imax(s, imin(*end, n));    /* clamp(end, start, n) search_exact.py:30 */
    *start = s;
    *end = e;
}

/* next occurrence wholly inside [from, end) or -1 (bytes.find semantics) */
static int64_t find_from(const uint8_t *sub, int64_t sublen, const uint8_t *seq, int64_t from,
                         int64_t end) {
    for (int64_t i = from; i + sublen <= end; i++) {
        if (seq[i] == sub[0] && memcmp(seq + i, sub, (size_t)sublen) == 0) return i;
    }
    return -1;
}

/* search_exact over the whole list; returns malloc'ed int64 array (caller fzo_free) */
FZO_API int64_t fzo_search_exact(const uint8_t *sub, int64_t sublen, const uint8_t *seq, int64_t n,
                                 int64_t start, int64_t end, int64_t **out) {
    *out = NULL;
    if (sublen <= 0) return -1; /* ValueError('subsequence must not be empty') :23-24 */
    clamp_window(n, &start, &end);
    int64_t cnt = 0, cap = 0;
    int64_t *v = NULL;
    int64_t idx = find_from(sub, sublen, seq, start, end);
    while (idx >= 0) {
        if (cnt == cap) {
            cap = cap ? cap * 2 : 64;
            v = (int64_t *)realloc(v, (size_t)cap * sizeof(int64_t));
        }
        v[cnt++] = idx;
        idx = find_from(sub, sublen, seq, idx + 1, end); /* :53-56 overlapping allowed */
    }
    *out = v;
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * _py_expand_short (levenshtein_ngram.py:22-74).  sub/seq are read with a stride so the
 * reversed slices of :186-188 need no copy.  Returns 1 and (*dist,*len) or 0 for (None,None).
 * ---------------------------------------------------------------------------------------- */
#define FZO_MAX_SUB 4096

static int expand_short(const uint8_t *sub, int64_t sstride, int sublen, const uint8_t *seq,
                        int64_t qstride, int seqlen, int max_l, int *dist, int *len) {
    if (sublen == 0) { /* :42-43 */
        *dist = 0;
        *len = 0;
        return 1;
    }
    int scores[FZO_MAX_SUB];
    for (int j = 0; j < sublen; j++) scores[j] = j + 1; /* :47 */
    int min_score = sublen;                             /* :49 */
    int min_score_idx = -1;                             /* :50 */
    for (int si = 0; si < seqlen; si++) {               /* :52 */
        uint8_t ch = seq[(int64_t)si * qstride];
        int a = si;    /* :54 */
        int c = a + 1; /* :55 */
        for (int j = 0; j < sublen; j++) { /* :56-63 */
            int b = scores[j];
            int v = a + (ch != sub[(int64_t)j * sstride]);
            if (b + 1 < v) v = b + 1;
            if (c + 1 < v) v = c + 1;
            c = scores[j] = v;
            a = b;
        }
        if (c <= min_score) { /* :66-68 */
            min_score = c;
            min_score_idx = si;
        } else { /* :71-72  elif min(scores) >= min_score: break */
            int mn = scores[0];
            for (int j = 1; j < sublen; j++)
                if (scores[j] < mn) mn = scores[j];
            if (mn >= min_score) break;
        }
    }
    if (min_score <= max_l) { /* :74 */
        *dist = min_score;
        *len = min_score_idx + 1;
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * _py_expand_long (levenshtein_ngram.py:77-143), band bookkeeping restated literally
 * (new_needle_idx_range_start is None  <=>  ns_none).
 * ---------------------------------------------------------------------------------------- */
static int expand_long(const uint8_t *sub, int64_t sstride, int sublen, const uint8_t *seq,
                       int64_t qstride, int seqlen, int max_l, int *dist, int *len) {
    if (sublen == 0) { /* :86-88 */
        *dist = 0;
        *len = 0;
        return 1;
    }
    int scores[FZO_MAX_SUB];
    for (int j = 0; j < sublen; j++) scores[j] = j + 1; /* :92 */
    int min_score = sublen;                             /* :94 */
    int min_score_idx = -1;
    int max_good_score = max_l; /* :96 */
    int new_start = 0;          /* :97 */
    int ns_none = 0;
    int new_end = sublen - 1; /* :98 */
    for (int si = 0; si < seqlen; si++) { /* :100 */
        uint8_t ch = seq[(int64_t)si * qstride];
        int rstart = new_start;                                 /* :102 */
        int rend = (int)imin(sublen, (int64_t)new_end + 1);     /* :103 */
        int a = si;                                             /* :105 */
        int c = a + 1;                                          /* :106 */
        if (c <= max_good_score) { /* :108-110 */
            new_start = 0;
            ns_none = 0;
            new_end = 0;
        } else { /* :111-113 */
            ns_none = 1;
            new_start = 0;
            new_end = -1;
        }
        for (int j = rstart; j < rend; j++) { /* :115-122 */
            int b = scores[j];
            int v = a + (ch != sub[(int64_t)j * sstride]);
            if (b + 1 < v) v = b + 1;
            if (c + 1 < v) v = c + 1;
            c = scores[j] = v;
            a = b;
            if (c <= max_good_score) { /* :124-130 */
                if (ns_none) {
                    ns_none = 0;
                    new_start = j;
                }
                int cand = j + 1 + (max_good_score - c);
                if (cand > new_end) new_end = cand;
            }
        }
        if (ns_none) break; /* :133-134 */
        if (rend == sublen && c <= min_score) { /* :137-141 */
            min_score = c;
            min_score_idx = si;
            if (min_score < max_good_score) max_good_score = min_score;
        }
    }
    if (min_score <= max_l) { /* :143 */
        *dist = min_score;
        *len = min_score_idx + 1;
        return 1;
    }
    return 0;
}

/* _expand dispatcher (levenshtein_ngram.py:8-19) */
static int expand(const uint8_t *sub, int64_t sstride, int sublen, const uint8_t *seq,
                  int64_t qstride, int seqlen, int max_l, int *dist, int *len) {
    int thr = max_l * 2 > 10 ? max_l * 2 : 10; /* :16 */
    if (sublen > thr) return expand_long(sub, sstride, sublen, seq, qstride, seqlen, max_l, dist, len);
    return expand_short(sub, sstride, sublen, seq, qstride, seqlen, max_l, dist, len);
}

/* exported for the known-answer vectors (tests/test_levenshtein.py:64-158 TestExpandBase).
 * which: 0 = _expand, 1 = _py_expand_short, 2 = _py_expand_long.  out[0]=dist out[1]=len. */
FZO_API int fzo_expand(int which, const uint8_t *sub, int sublen, const uint8_t *seq, int seqlen,
                       int max_l, int *out) {
    if (sublen > FZO_MAX_SUB) return -1;
    int d = 0, l = 0, r;
    if (which == 1)
        r = expand_short(sub, 1, sublen, seq, 1, seqlen, max_l, &d, &l);
    else if (which == 2)
        r = expand_long(sub, 1, sublen, seq, 1, seqlen, max_l, &d, &l);
    else
        r = expand(sub, 1, sublen, seq, 1, seqlen, max_l, &d, &l);
    out[0] = d;
    out[1] = l;
    return r;
}

/* ------------------------------------------------------------------------------------------
 * find_near_matches_levenshtein_ngrams (levenshtein_ngram.py:159-198): the raw match stream.
 * Returns count (>=0), -2 for the ValueError at :163-165.  If ngram_ids/idxs are non-NULL
 * they receive, per raw match, the n-gram ordinal and the n-gram hit index (for L0 parity
 * keyed by (ngram, idx)).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t *v;
    int64_t n, cap;
} i64vec;
static void i64push(i64vec *V, int64_t x) {
    if (V->n == V->cap) {
        V->cap = V->cap ? V->cap * 2 : 64;
        V->v = (int64_t *)realloc(V->v, (size_t)V->cap * sizeof(int64_t));
    }
    V->v[V->n++] = x;
}

FZO_API int64_t fzo_levenshtein_ngrams(const uint8_t *P, int m, const uint8_t *H, int64_t n, int k,
                                       fzo_match **out, int64_t **out_ngram, int64_t **out_idx) {
    *out = NULL;
    if (out_ngram) *out_ngram = NULL;
    if (out_idx) *out_idx = NULL;
    if (m > FZO_MAX_SUB) return -1;
    int L = m / (k + 1); /* :162 */
    if (L == 0) return -2; /* :163-165 */
    fzo_list R = {0, 0, 0};
    i64vec NG = {0, 0, 0}, IX = {0, 0, 0};
    int ord = 0;
    for (int s = 0; s <= m - L; s += L, ord++) { /* :170 range(0, m-L+1, L) */
        int ngram_end = s + L;                   /* :171 */
        int64_t start_index = imax(0, (int64_t)s - k);                        /* :174 */
        int64_t end_index = imin(n, n - m + ngram_end + k);                   /* :175 */
        int64_t ws = start_index, we = end_index;
        clamp_window(n, &ws, &we);
        int64_t idx = find_from(P + s, L, H, ws, we); /* :176 */
        while (idx >= 0) {
            int64_t p0 = idx - s;
            /* right: _expand(P[ngram_end:], H[idx+L : p0+m+k], k)  :178-182 */
            int64_t rlo = imin(idx + L, n);
            int64_t rhi = imin(imax(p0 + m + k, 0), n);
            if (rhi < rlo) rhi = rlo;
            int dr, rs;
            if (expand(P + ngram_end, 1, m - ngram_end, H + rlo, 1, (int)(rhi - rlo), k, &dr, &rs)) {
                /* left: _expand(P[:s][::-1], H[max(0,p0-(k-dr)) : idx][::-1], k-dr)  :185-189 */
                int64_t llo = imax(0, p0 - (k - dr));
                int64_t lhi = idx;
                if (llo > lhi) llo = lhi;
                int dl, ls;
                if (expand(P + s - 1, -1, s, H + lhi - 1, -1, (int)(lhi - llo), k - dr, &dl, &ls)) {
                    push(&R, idx - ls, idx + L + rs, dl + dr); /* :194-198 */
                    i64push(&NG, ord);
                    i64push(&IX, idx);
                }
            }
            idx = find_from(P + s, L, H, idx + 1, we);
        }
    }
    *out = R.v;
    if (out_ngram) *out_ngram = NG.v; else free(NG.v);
    if (out_idx) *out_idx = IX.v; else free(IX.v);
    return R.n;
}

/* ------------------------------------------------------------------------------------------
 * find_near_matches_levenshtein_linear_programming (levenshtein.py:52-148), literal.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t start;
    int32_t j, d;
} lp_cand;

typedef struct {
    lp_cand *v;
    int64_t n, cap;
} lp_vec;
static void lp_push(lp_vec *V, int64_t start, int j, int d) {
    if (V->n == V->cap) {
        V->cap = V->cap ? V->cap * 2 : 64;
        V->v = (lp_cand *)realloc(V->v, (size_t)V->cap * sizeof(lp_cand));
    }
    V->v[V->n].start = start;
    V->v[V->n].j = j;
    V->v[V->n].d = d;
    V->n++;
}

FZO_API int64_t fzo_levenshtein_lp(const uint8_t *P, int m, const uint8_t *H, int64_t n, int k,
                                   fzo_match **out) {
    *out = NULL;
    if (m <= 0) return -1; /* ValueError :54-55 */
    fzo_list R = {0, 0, 0};
    if (k >= m) { /* :62-65 */
        for (int64_t i = 0; i <= n; i++) push(&R, i, i, m);
        *out = R.v;
        return R.n;
    }
    /* make_char2first_subseq_index :44-49 -- first index of each char within P[:k+1] */
    int first[256];
    for (int c = 0; c < 256; c++) first[c] = -1;
    for (int j = imin(k, m - 1); j >= 0; j--) first[P[j]] = j;

    lp_vec cur = {0, 0, 0}, nxt = {0, 0, 0};
    for (int64_t index = 0; index < n; index++) { /* :73 */
        uint8_t ch = H[index];
        nxt.n = 0;
        int j0 = first[ch]; /* :76 */
        if (j0 >= 0) {
            if (j0 + 1 == m) /* :78-79 */
                push(&R, index, index + 1, j0);
            else
                lp_push(&nxt, index, j0 + 1, j0); /* :80-81 */
        }
        for (int64_t ci = 0; ci < cur.n; ci++) { /* :83 */
            lp_cand cand = cur.v[ci];
            if (P[cand.j] == ch) { /* :85 */
                if (cand.j + 1 == m)
                    push(&R, cand.start, index + 1, cand.d); /* :87-88 */
                else
                    lp_push(&nxt, cand.start, cand.j + 1, cand.d); /* :90-93 */
            } else {
                if (cand.d == k) continue;               /* :100-101 */
                lp_push(&nxt, cand.start, cand.j, cand.d + 1); /* :104 */
                if (index + 1 < n && cand.j + 1 < m)           /* :106 */
                    lp_push(&nxt, cand.start, cand.j + 1, cand.d + 1); /* :109-112 */
                for (int t = 1; t <= k - cand.d; t++) { /* :115 */
                    if (cand.j + t == m) {              /* :118 */
                        push(&R, cand.start, index + 1, cand.d + t);
                        break;
                    } else if (P[cand.j + t] == ch) { /* :126 */
                        if (cand.j + t + 1 == m)      /* :129 */
                            push(&R, cand.start, index + 1, cand.d + t);
                        else
                            lp_push(&nxt, cand.start, cand.j + 1 + t, cand.d + t); /* :135-138 */
                        break;
                    }
                }
            }
        }
        lp_vec tmp = cur; /* :143 */
        cur = nxt;
        nxt = tmp;
    }
    for (int64_t ci = 0; ci < cur.n; ci++) { /* :145-148 */
        int d = cur.v[ci].d + m - cur.v[ci].j;
        if (d <= k) push(&R, cur.v[ci].start, n, d);
    }
    free(cur.v);
    free(nxt.v);
    *out = R.v;
    return R.n;
}

/* ------------------------------------------------------------------------------------------
 * _find_near_matches_generic_linear_programming (generic_search.py:57-177), literal.
 * Appends to R with `shift` added to start/end (used by the n-gram driver, :230-237).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t start;
    int32_t j, l, ns, ni, nd;
} g_cand;
typedef struct {
    g_cand *v;
    int64_t n, cap;
} g_vec;
static void g_push(g_vec *V, g_cand c) {
    if (V->n == V->cap) {
        V->cap = V->cap ? V->cap * 2 : 64;
        V->v = (g_cand *)realloc(V->v, (size_t)V->cap * sizeof(g_cand));
    }
    V->v[V->n++] = c;
}

static void generic_lp(const uint8_t *P, int m, const uint8_t *H, int64_t n, int max_subs,
                       int max_ins, int max_dels, int max_l, int64_t shift, fzo_list *R) {
    g_vec cur = {0, 0, 0}, nxt = {0, 0, 0};
    for (int64_t index = 0; index < n; index++) { /* :79 */
        uint8_t ch = H[index];
        g_cand fresh = {index, 0, 0, 0, 0, 0};
        g_push(&cur, fresh); /* :81 */
        nxt.n = 0;
        for (int64_t ci = 0; ci < cur.n; ci++) { /* :84 */
            g_cand cand = cur.v[ci];
            if (ch == P[cand.j]) { /* :86 */
                if (cand.j + 1 == m)
                    push(R, cand.start + shift, index + 1 + shift, cand.l); /* :88-89 */
                else {
                    g_cand c2 = cand; /* :91-94 */
                    c2.j++;
                    g_push(&nxt, c2);
                }
            } else {
                if (cand.l == max_l) continue; /* :101-102 */
                if (cand.ni < max_ins) {       /* :104-109 */
                    g_cand c2 = cand;
                    c2.ni++;
                    c2.l++;
                    g_push(&nxt, c2);
                }
                if (cand.j + 1 < m) {         /* :111 */
                    if (cand.ns < max_subs) { /* :112-119 */
                        g_cand c2 = cand;
                        c2.ns++;
                        c2.j++;
                        c2.l++;
                        g_push(&nxt, c2);
                    } else if (cand.nd < max_dels && cand.ni < max_ins) { /* :120-128 */
                        g_cand c2 = cand;
                        c2.ni++;
                        c2.nd++;
                        c2.j++;
                        c2.l++;
                        g_push(&nxt, c2);
                    }
                } else { /* :129-138 */
                    if (cand.ns < max_subs || (cand.nd < max_dels && cand.ni < max_ins))
                        push(R, cand.start + shift, index + 1 + shift, cand.l + 1);
                }
                int lim = (int)imin(max_dels - cand.nd, max_l - cand.l); /* :141 */
                for (int t = 1; t <= lim; t++) {
                    if (cand.j + t == m) { /* :144-147 */
                        push(R, cand.start + shift, index + shift, cand.l + t);
                        break;
                    } else if (P[cand.j + t] == ch) { /* :151 */
                        if (cand.j + t + 1 == m)      /* :154-156 */
                            push(R, cand.start + shift, index + shift, cand.l + t);
                        else { /* :159-164 */
                            g_cand c2 = cand;
                            c2.nd += t;
                            c2.j += 1 + t;
                            c2.l += t;
                            g_push(&nxt, c2);
                        }
                        break;
                    }
                }
            }
        }
        g_vec tmp = cur; /* :170 */
        cur = nxt;
        nxt = tmp;
    }
    for (int64_t ci = 0; ci < cur.n; ci++) { /* :172-177 */
        g_cand cand = cur.v[ci];
        int t = m - cand.j;
        if (cand.nd + t <= max_dels && cand.l + t <= max_l)
            push(R, cand.start + shift, n + shift, cand.l + t);
    }
    free(cur.v);
    free(nxt.v);
}

FZO_API int64_t fzo_generic_lp(const uint8_t *P, int m, const uint8_t *H, int64_t n, int max_subs,
                               int max_ins, int max_dels, int max_l, fzo_match **out) {
    *out = NULL;
    if (m <= 0) return -1; /* :67-68 */
    fzo_list R = {0, 0, 0};
    generic_lp(P, m, H, n, max_subs, max_ins, max_dels, max_l, 0, &R);
    *out = R.v;
    return R.n;
}

/* find_near_matches_generic_ngrams (generic_search.py:198-237) */
FZO_API int64_t fzo_generic_ngrams(const uint8_t *P, int m, const uint8_t *H, int64_t n,
                                   int max_subs, int max_ins, int max_dels, int max_l,
                                   fzo_match **out) {
    *out = NULL;
    if (m <= 0) return -1; /* :208-209 */
    int k = max_l;
    int L = m / (k + 1); /* :217 */
    if (L == 0) return -2; /* :218-219 */
    fzo_list R = {0, 0, 0};
    for (int s = 0; s <= m - L; s += L) { /* :221 */
        int ngram_end = s + L;
        int64_t start_index = imax(0, (int64_t)s - k);      /* :223 */
        int64_t end_index = imin(n, n - m + ngram_end + k); /* :224 */
        if (end_index <= start_index) continue;             /* :225-226 */
        int64_t ws = start_index, we = end_index;
        clamp_window(n, &ws, &we);
        int64_t idx = find_from(P + s, L, H, ws, we); /* :227 */
        while (idx >= 0) {
            int64_t p0 = idx - s;
            int64_t lo = imax(0, p0 - k);             /* :231,234 */
            int64_t hi = imin(n, imax(0, p0 + m + k)); /* slice end clipped like Python */
            if (hi < lo) hi = lo;
            generic_lp(P, m, H + lo, hi - lo, max_subs, max_ins, max_dels, max_l, lo, &R);
            idx = find_from(P + s, L, H, idx + 1, we);
        }
    }
    *out = R.v;
    return R.n;
}

/* ------------------------------------------------------------------------------------------
 * substitutions only.
 *   _find_near_matches_substitutions_lp      (substitutions_only.py:82-136)
 *   find_near_matches_substitutions_ngrams   (substitutions_only.py:148-215), incl. the
 *   first-seen de-duplication by start and the sort by start of :160-167.
 * ---------------------------------------------------------------------------------------- */
FZO_API int64_t fzo_subs_lp(const uint8_t *P, int m, const uint8_t *H, int64_t n, int k,
                            fzo_match **out) {
    *out = NULL;
    if (m <= 0) return -1;
    fzo_list R = {0, 0, 0};
    /* ring of per-alignment match counters (:104-136); ring[(head+i)%m] == candidates[i] */
    int64_t *ring = (int64_t *)calloc((size_t)m, sizeof(int64_t));
    int64_t head = 0;   /* index of candidates[0] */
    int64_t filled = 1; /* deque([0], maxlen=m) :104 */
    int64_t index = 0;
    for (; index < n && index < m - 1; index++) { /* :105-108 */
        uint8_t ch = H[index];
        for (int j = 0; j < m; j++)
            if (P[j] == ch && j <= index) ring[(head + j) % m] += 1; /* :106-107 */
        head = (head - 1 + m) % m; /* appendleft(0) :108 */
        ring[head] = 0;
        if (filled < m) filled++;
    }
    for (; index < n; index++) { /* :114 */
        uint8_t ch = H[index];
        for (int j = 0; j < m; j++)
            if (P[j] == ch) ring[(head + j) % m] += 1; /* :115-116 */
        head = (head - 1 + m) % m;                     /* rotate(1) :119 */
        int64_t n_subs = m - ring[head];               /* :121 */
        ring[head] = 0;                                /* :123 */
        if (n_subs <= k) push(&R, index - (m - 1), index + 1, n_subs); /* :126-131 */
    }
    free(ring);
    *out = R.v;
    return R.n;
}

/* count_differences_with_maximum (common.py:119-126) */
static int count_diff_max(const uint8_t *a, const uint8_t *b, int64_t len, int maxd) {
    int nd = 0;
    for (int64_t i = 0; i < len; i++) {
        if (a[i] != b[i]) {
            nd++;
            if (nd == maxd) return nd;
        }
    }
    return nd;
}

typedef struct {
    fzo_match mt;
    int64_t ord;
} fzo_tagged;

/* order by (start, generation ordinal): a stable sort by start */
static int cmp_tagged(const void *a, const void *b) {
    const fzo_tagged *x = (const fzo_tagged *)a, *y = (const fzo_tagged *)b;
    if (x->mt.start != y->mt.start) return x->mt.start < y->mt.start ? -1 : 1;
    return x->ord < y->ord ? -1 : (x->ord > y->ord);
}

FZO_API int64_t fzo_subs_ngrams(const uint8_t *P, int m, const uint8_t *H, int64_t n, int k,
                                fzo_match **out) {
    *out = NULL;
    if (m <= 0) return -1;
    int L = m / (k + 1); /* :178 */
    if (L == 0) return -2; /* :179-182 */
    fzo_list R = {0, 0, 0};
    for (int s = 0; s <= m - L; s += L) { /* :184 */
        int ngram_end = s + L;
        int64_t ws = s, we = n - (m - ngram_end); /* :190 */
        clamp_window(n, &ws, &we);
        int64_t idx = find_from(P + s, L, H, ws, we);
        while (idx >= 0) {
            int nsub = 0;
            int skip = 0;
            /* seq_before = H[idx-s : idx]  :193-199 */
            if (s > 0 && memcmp(P, H + idx - s, (size_t)s) != 0) {
                nsub += count_diff_max(H + idx - s, P, s, k - nsub + 1);
                if (nsub > k) skip = 1;
            }
            if (!skip) {
                int64_t alen = m - ngram_end; /* seq_after :201 */
                if (alen > 0 && memcmp(P + ngram_end, H + idx + L, (size_t)alen) != 0) {
                    if (nsub == k)
                        skip = 1; /* :203-204 */
                    else {
                        nsub += count_diff_max(H + idx + L, P + ngram_end, alen, k - nsub + 1);
                        if (nsub > k) skip = 1; /* :208-209 */
                    }
                }
            }
            if (!skip) push(&R, idx - s, idx - s + m, nsub); /* :211-215 */
            idx = find_from(P + s, L, H, idx + 1, we);
        }
    }
    /* :160-167 keep the first match per start (generation order), then sort by start */
    if (R.n > 1) {
        fzo_tagged *T = (fzo_tagged *)malloc((size_t)R.n * sizeof(fzo_tagged));
        for (int64_t i = 0; i < R.n; i++) {
            T[i].mt = R.v[i];
            T[i].ord = i;
        }
        qsort(T, (size_t)R.n, sizeof(fzo_tagged), cmp_tagged);
        int64_t w = 0;
        for (int64_t i = 0; i < R.n; i++) {
            if (w > 0 && R.v[w - 1].start == T[i].mt.start) continue;
            R.v[w++] = T[i].mt;
        }
        R.n = w;
        free(T);
    }
    *out = R.v;
    return R.n;
}

/* ------------------------------------------------------------------------------------------
 * consolidate_overlapping_matches (common.py:145-189), literal group_matches; the winner of
 * a group is min by (dist, -(end-start)) (:180-182); the reference breaks ties by set order
 * (hash-seed dependent) -- here ties go to the smallest (start, end), which is the
 * product's documented deterministic rule.  Output sorted by (start, end, dist) (:189).
 * group_of[i] (optional) receives the group ordinal of raw match i.
 * ---------------------------------------------------------------------------------------- */
static int cmp_match_full(const void *a, const void *b) {
    const fzo_match *x = (const fzo_match *)a, *y = (const fzo_match *)b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return 0;
}

FZO_API int64_t fzo_consolidate(const fzo_match *raw, int64_t n, fzo_match **out,
                                int64_t *group_of) {
    *out = NULL;
    if (n == 0) return 0;
    /* groups as disjoint-set over matches, with (start,end) hull per live group */
    int64_t *gid = (int64_t *)malloc((size_t)n * sizeof(int64_t)); /* match -> group id */
    int64_t *gstart = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t *gend = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    char *alive = (char *)calloc((size_t)n, 1);
    int64_t ng = 0;
    for (int64_t i = 0; i < n; i++) { /* :163 */
        int64_t first = -1;
        int64_t nover = 0;
        for (int64_t g = 0; g < ng; g++) { /* :164 is_match_in_group :152-153 */
            if (!alive[g]) continue;
            if (!(raw[i].end <= gstart[g] || raw[i].start >= gend[g])) {
                if (first < 0) first = g;
                nover++;
            }
        }
        if (nover == 0) { /* :165-166 */
            gstart[ng] = raw[i].start;
            gend[ng] = raw[i].end;
            alive[ng] = 1;
            gid[i] = ng++;
        } else if (nover == 1) { /* :167-168 add_match :155-158 */
            gid[i] = first;
            gstart[first] = imin(gstart[first], raw[i].start);
            gend[first] = imax(gend[first], raw[i].end);
        } else { /* :169-175 merge all overlapping groups into a new one */
            int64_t ns = raw[i].start, ne = raw[i].end;
            int64_t newg = ng++;
            /* decide membership against the pre-merge hulls first */
            char *hit = (char *)calloc((size_t)newg, 1);
            for (int64_t g = 0; g < newg; g++) {
                if (!alive[g]) continue;
                if (!(raw[i].end <= gstart[g] || raw[i].start >= gend[g])) hit[g] = 1;
            }
            for (int64_t g = 0; g < newg; g++) {
                if (!hit[g]) continue;
                ns = imin(ns, gstart[g]);
                ne = imax(ne, gend[g]);
                alive[g] = 0;
                for (int64_t q = 0; q < i; q++)
                    if (gid[q] == g) gid[q] = newg;
            }
            free(hit);
            gstart[newg] = ns;
            gend[newg] = ne;
            alive[newg] = 1;
            gid[i] = newg;
        }
    }
    /* best per group :180-189 */
    fzo_match *best = (fzo_match *)malloc((size_t)ng * sizeof(fzo_match));
    char *has = (char *)calloc((size_t)ng, 1);
    for (int64_t i = 0; i < n; i++) {
        int64_t g = gid[i];
        if (!has[g]) {
            best[g] = raw[i];
            has[g] = 1;
            continue;
        }
        fzo_match *b = &best[g];
        int64_t li = raw[i].end - raw[i].start, lb = b->end - b->start;
        int better = 0;
        if (raw[i].dist != b->dist)
            better = raw[i].dist < b->dist;
        else if (li != lb)
            better = li > lb;
        else if (raw[i].start != b->start)
            better = raw[i].start < b->start;
        else
            better = raw[i].end < b->end;
        if (better) *b = raw[i];
    }
    fzo_match *res = (fzo_match *)malloc((size_t)ng * sizeof(fzo_match));
    int64_t w = 0;
    for (int64_t g = 0; g < ng; g++)
        if (alive[g] && has[g]) res[w++] = best[g];
    qsort(res, (size_t)w, sizeof(fzo_match), cmp_match_full);
    if (group_of)
        for (int64_t i = 0; i < n; i++) group_of[i] = gid[i];
    free(gid);
    free(gstart);
    free(gend);
    free(alive);
    free(best);
    free(has);
    *out = res;
    return w;
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic benchmark corpus (NOT a restatement of the reference, which has no corpus generator:
 * SURVEY F9/F10).  Counter-based: bytes 4q..4q+3 of the global sequence come from one splitmix64
 * hash of (seed, q), 16 bits per byte -- the same definition as the device generator
 * (fuzzysearch_b200/csrc/common.cuh: synth_word), restated here so that bench.py's reference arm
 * and the parity checks can rebuild any slice of the corpus WITHOUT loading the product library.
 * ------------------------------------------------------------------------------------------------ */
static uint64_t fzo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

FZO_API void fzo_synth(uint8_t *dst, uint64_t global_offset, uint64_t n, const uint8_t *alphabet,
                       uint32_t alen, uint64_t seed) {
    uint64_t i = 0;
    while (i < n) {
        uint64_t g = global_offset + i, q = g / 4;
        uint64_t x = fzo_splitmix64(seed ^ (q * 0xD1342543DE82EF95ull));
        for (uint64_t b = g % 4; b < 4 && i < n; b++, i++) {
            uint32_t r = (uint32_t)(x >> (16 * b)) & 0xFFFFu;
            dst[i] = alphabet[(r * alen) >> 16];
        }
    }
}
