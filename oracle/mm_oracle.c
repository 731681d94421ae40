/*
 * mm_oracle.c -- CPU restatement of Monkey-Moore's relative-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see mm_oracle.h).  Plain C, no dependencies.
 * The algorithm is restated as a pure per-window function F(s) -> (match,
 * jump) plus the sequential chain  s <- s + jump(s)  that defines which
 * windows the reference actually visits (its Boyer-Moore-style skip is lossy,
 * so the result set is defined by that chain, not by "all matching windows").
 *
 * Citations are file:line under /root/reference/.
 */
#include "mm_oracle.h"

#include <stdlib.h>
#include <string.h>

struct mmo_pattern {
    int bits;            /* 8 or 16 */
    int64_t vmax;        /* std::numeric_limits<Ty>::max() */
    int mode;            /* 0 simple, 1 wildcard, 2 value scan */
    int L;
    uint32_t *kw;        /* keyword as given */
    uint32_t *nk;        /* case-normalised keyword (wildcard mode) */
    uint32_t wildcard;
    int nseq;
    uint32_t *seq;       /* custom character sequence */
    /* char -> index map (std::map<CharType,int> custom_character_index),
     * including the default-inserted zeros that operator[] creates for
     * characters that are not part of the sequence. */
    int nidx, capidx;
    uint32_t *idx_key;
    int *idx_val;

    int has_case_change;
    int mostly_lowercase;

    int *ed;             /* expected_diff[i] */
    int *prev;           /* index compared against at position i (bridged) */
    unsigned char *lit;  /* is_literal_map */
    unsigned char *wst;  /* wildcard_skip_table */

    /* sparse skip table: value for diff tab_key[j] is tab_val[j]; every other
     * diff maps to tab_default.  Values are the raw table entries (before the
     * max(.,1) the scan applies). */
    int ntab;
    int *tab_key;
    int *tab_val;
    int tab_default;

    int lead;            /* leading wildcards  (count_prefix_length) */
    int first_lit;       /* first_non_wildcard_index */
    int opp_idx;         /* first opposite-case letter in keyword, or -1 */
};

/* include/mmoore/text_utils.hpp:41-50 -- (c < 128) && isupper/islower in the
 * "C" locale. */
static int ascii_upper(uint32_t c) { return c >= 'A' && c <= 'Z'; }
static int ascii_lower(uint32_t c) { return c >= 'a' && c <= 'z'; }

/* std::map<CharType,int>::operator[] semantics: unknown key is inserted with
 * value 0 (src/core/monkey_moore.cpp:239-240, 387, 516, 578-581). */
static int seq_index(mmo_pattern *p, uint32_t c) {
    for (int i = 0; i < p->nidx; i++)
        if (p->idx_key[i] == c) return p->idx_val[i];
    if (p->nidx == p->capidx) {
        p->capidx = p->capidx ? p->capidx * 2 : 64;
        p->idx_key = (uint32_t *)realloc(p->idx_key, sizeof(uint32_t) * p->capidx);
        p->idx_val = (int *)realloc(p->idx_val, sizeof(int) * p->capidx);
    }
    p->idx_key[p->nidx] = c;
    p->idx_val[p->nidx] = 0;
    p->nidx++;
    return 0;
}

static void seq_assign(mmo_pattern *p, uint32_t c, int v) {
    for (int i = 0; i < p->nidx; i++)
        if (p->idx_key[i] == c) { p->idx_val[i] = v; return; }
    seq_index(p, c);
    p->idx_val[p->nidx - 1] = v;
}

/* value of a keyword character: code point, or index in the sequence */
static int64_t char_value(mmo_pattern *p, uint32_t c) {
    if (p->nseq > 0) return seq_index(p, c);
    return (int64_t)c;
}

static void tab_set(mmo_pattern *p, int key, int val, int overwrite) {
    for (int j = 0; j < p->ntab; j++) {
        if (p->tab_key[j] == key) {
            if (overwrite) p->tab_val[j] = val;
            return;
        }
    }
    p->tab_key[p->ntab] = key;
    p->tab_val[p->ntab] = val;
    p->ntab++;
}

static int tab_get(const mmo_pattern *p, int key) {
    for (int j = 0; j < p->ntab; j++)
        if (p->tab_key[j] == key) return p->tab_val[j];
    return p->tab_default;
}

void mmo_free(mmo_pattern *p) {
    if (!p) return;
    free(p->kw); free(p->nk); free(p->seq); free(p->idx_key); free(p->idx_val);
    free(p->ed); free(p->prev); free(p->lit); free(p->wst);
    free(p->tab_key); free(p->tab_val);
    free(p);
}

/* C's conversion of a keyword difference to `int`: the reference subtracts
 * char32_t (uint32_t) values and stores the result in an int
 * (src/core/monkey_moore.cpp:236, 560-563), i.e. arithmetic modulo 2^32. */
static int diff32(int64_t a, int64_t b) {
    return (int)(int32_t)(uint32_t)((uint64_t)a - (uint64_t)b);
}

static int in_table_range(const mmo_pattern *p, int d) {
    /* src/core/monkey_moore.cpp:128-130, 259-261:
     * index = diff + max();  0 <= index < 2*(max()+1) */
    int64_t index = (int64_t)d + p->vmax;
    return index >= 0 && index < 2 * (p->vmax + 1);
}

/* src/core/monkey_moore.cpp:106-142 */
static int preprocess_simple(mmo_pattern *p) {
    int L = p->L;
    for (int i = 0; i < L; i++) {
        p->lit[i] = 1;
        p->prev[i] = i == 0 ? L - 1 : i - 1;
        p->ed[i] = diff32(char_value(p, p->kw[i]), char_value(p, p->kw[p->prev[i]]));
    }
    /* compute_relative_values evaluates [0] first then L-1..1; with the
     * default-inserting map the order of evaluation never changes a value
     * (unknown characters always read 0), so a plain loop is equivalent
     * (src/core/monkey_moore.cpp:551-585). */
    p->tab_default = L - 1;
    for (int i = L - 1; i >= 0; i--) {
        if (!in_table_range(p, p->ed[i])) return MMO_ERR_SKIP_OOB;
        /* first writer wins: a slot is only written while it still holds the
         * default L-1 (src/core/monkey_moore.cpp:133-135) */
        tab_set(p, p->ed[i], L - i - 1, 0);
    }
    p->lead = 0;
    p->first_lit = 0;
    p->opp_idx = -1;
    return MMO_OK;
}

/* src/core/monkey_moore.cpp:144-304 */
static int preprocess_wildcard(mmo_pattern *p) {
    int L = p->L;
    memcpy(p->nk, p->kw, sizeof(uint32_t) * L);

    if (p->nseq == 0) {
        int up = 0, lo = 0;
        for (int i = 0; i < L; i++) { up += ascii_upper(p->kw[i]); lo += ascii_lower(p->kw[i]); }
        p->mostly_lowercase = lo > up;                       /* :163 */
        if (up > 0 && lo > 0) {                              /* :165-180 */
            for (int i = 0; i < L; i++) {
                if (up > lo) { if (ascii_lower(p->nk[i])) p->nk[i] = p->wildcard; }
                else         { if (ascii_upper(p->nk[i])) p->nk[i] = p->wildcard; }
            }
        }
    }

    int nlit = 0, first = -1, last = -1;
    for (int i = 0; i < L; i++) {
        p->lit[i] = p->nk[i] != p->wildcard;                 /* :190-197 */
        p->ed[i] = 0;
        p->prev[i] = i;                                      /* bridge offset 0 */
        if (p->lit[i]) { if (first < 0) first = i; last = i; nlit++; }
    }
    /* :222-247 -- bridge every literal to the previous literal, the first one
     * wraps around to the last */
    int prevlit = last;
    for (int i = 0; i < L; i++) {
        if (!p->lit[i]) continue;
        p->prev[i] = prevlit;
        p->ed[i] = diff32(char_value(p, p->nk[i]), char_value(p, p->nk[prevlit]));
        prevlit = i;
    }

    /* :249-276 -- skip table, written through a `char` cast; later (smaller i)
     * writes override earlier ones; i = 0 is not visited */
    p->tab_default = (int)(signed char)(L - 1);
    for (int i = L - 1; i > 0; --i) {
        if (!in_table_range(p, p->ed[i])) return MMO_ERR_SKIP_OOB;
        int wc_after = 0;
        for (int j = i + 1; j < L; j++) wc_after += p->nk[j] == p->wildcard;
        tab_set(p, p->ed[i], (int)(signed char)(L - wc_after - i - 1), 1);
    }

    /* :278-303 */
    for (int i = L - 1; i >= 0; --i) {
        if (p->nk[i] == p->wildcard) { p->wst[i] = 1; continue; }
        int lw = -1;
        for (int j = 0; j < i; j++) if (p->nk[j] == p->wildcard) lw = j;   /* find_last_index */
        if (lw == -1) lw = 0;
        int v = i - lw - 1; if (v < 1) v = 1;
        p->wst[i] = (unsigned char)v;
    }

    /* :438-447 */
    p->lead = 0;
    while (p->lead < L && p->nk[p->lead] == p->wildcard) p->lead++;
    p->first_lit = first < 0 ? L : first;

    /* :490-499 */
    p->opp_idx = -1;
    if (p->has_case_change) {
        for (int i = 0; i < L; i++) {
            int t = p->mostly_lowercase ? ascii_upper(p->kw[i]) : ascii_lower(p->kw[i]);
            if (t) { p->opp_idx = i; break; }
        }
    }
    (void)nlit;
    return MMO_OK;
}

static mmo_pattern *compile_common(const uint32_t *kw, int L, uint32_t wildcard,
                                   const uint32_t *seq, int nseq, int bits,
                                   int value_scan, int *err) {
    *err = MMO_OK;
    if (bits != 8 && bits != 16) { *err = MMO_ERR_ARG; return NULL; }
    if (L <= 0) { *err = MMO_ERR_EMPTY; return NULL; }

    mmo_pattern *p = (mmo_pattern *)calloc(1, sizeof(*p));
    p->bits = bits;
    p->vmax = bits == 8 ? 255 : 65535;
    p->L = L;
    p->wildcard = wildcard;
    p->kw = (uint32_t *)malloc(sizeof(uint32_t) * L);
    p->nk = (uint32_t *)malloc(sizeof(uint32_t) * L);
    memcpy(p->kw, kw, sizeof(uint32_t) * L);
    memcpy(p->nk, kw, sizeof(uint32_t) * L);
    p->nseq = nseq;
    p->seq = (uint32_t *)malloc(sizeof(uint32_t) * (nseq > 0 ? nseq : 1));
    if (nseq > 0) memcpy(p->seq, seq, sizeof(uint32_t) * nseq);
    p->ed = (int *)calloc(L, sizeof(int));
    p->prev = (int *)calloc(L, sizeof(int));
    p->lit = (unsigned char *)calloc(L, 1);
    p->wst = (unsigned char *)calloc(L, 1);
    p->tab_key = (int *)calloc(L + 1, sizeof(int));
    p->tab_val = (int *)calloc(L + 1, sizeof(int));

    /* initialize(): src/core/monkey_moore.cpp:54-78 */
    int has_wc = 0;
    for (int i = 0; i < L; i++) has_wc |= kw[i] == wildcard;
    p->has_case_change = 0;
    if (nseq == 0 && !value_scan) {
        int up = 0, lo = 0;
        for (int i = 0; i < L; i++) { up += ascii_upper(kw[i]); lo += ascii_lower(kw[i]); }
        p->has_case_change = up > 0 && lo > 0;
    }
    if (value_scan) p->mode = 2;
    else p->mode = (has_wc || p->has_case_change) ? 1 : 0;

    /* preprocess(): :83-100 -- later duplicates in the sequence win */
    for (int i = 0; i < nseq; i++) seq_assign(p, seq[i], i);

    int rc = p->mode == 1 ? preprocess_wildcard(p) : preprocess_simple(p);
    if (rc != MMO_OK) { *err = rc; mmo_free(p); return NULL; }

    /* The reference advances by L-1-lead after a match
     * (src/core/monkey_moore.cpp:398, 526); a non-positive advance never
     * terminates.  The GUI prevents it (src/gui/monkey_frame.cpp:1040,1099);
     * this restatement (and the product) reject it. */
    if (L - 1 - p->lead < 1) { *err = MMO_ERR_HANG; mmo_free(p); return NULL; }
    return p;
}

mmo_pattern *mmo_compile_keyword(const uint32_t *keyword, int keyword_len,
                                 uint32_t wildcard, const uint32_t *char_seq,
                                 int char_seq_len, int bits, int *err) {
    return compile_common(keyword, keyword_len, wildcard, char_seq, char_seq_len, bits, 0, err);
}

mmo_pattern *mmo_compile_values(const int16_t *values, int n, int bits, int *err) {
    *err = MMO_OK;
    if (n <= 0) { *err = MMO_ERR_EMPTY; return NULL; }
    uint32_t *kw = (uint32_t *)malloc(sizeof(uint32_t) * n);
    /* static_cast<CharType>(short): src/core/monkey_moore.cpp:33-35 */
    for (int i = 0; i < n; i++) kw[i] = (uint32_t)(int32_t)values[i];
    mmo_pattern *p = compile_common(kw, n, 0, NULL, 0, bits, 1, err);
    free(kw);
    return p;
}

int mmo_keyword_len(const mmo_pattern *p) { return p->L; }
int mmo_mode(const mmo_pattern *p) { return p->mode; }
int mmo_elem_bits(const mmo_pattern *p) { return p->bits; }

static inline uint32_t load_elem(const mmo_pattern *p, const uint8_t *base, uint64_t i) {
    if (p->bits == 8) return base[i];
    uint16_t v;
    memcpy(&v, base + 2 * i, 2);   /* host order, possibly unaligned */
    return v;
}

/* F(s): src/core/monkey_moore.cpp:347-405 (simple / value scan) and :449-541
 * (wildcard).  Returns 1 on match; *jump receives the advance. */
static int window(const mmo_pattern *p, const uint8_t *d, uint64_t s, int *jump) {
    int L = p->L;
    if (p->mode != 1) {
        /* exact signed differences, right to left; index 0 compares against
         * L-1 and can never be the first failure */
        for (int i = L - 1; i >= 0; i--) {
            int diff = (int)load_elem(p, d, s + i) - (int)load_elem(p, d, s + p->prev[i]);
            if (diff != p->ed[i]) {
                int sk = tab_get(p, diff);
                *jump = sk > 1 ? sk : 1;                       /* :403 */
                return 0;
            }
        }
        *jump = L - 1;                                         /* :398 */
        return 1;
    }
    uint32_t mask = (uint32_t)p->vmax;
    for (int i = L - 1; i >= 0; i--) {
        if (!p->lit[i]) continue;                              /* mask 0, expected 0 */
        uint32_t cur = load_elem(p, d, s + i), prv = load_elem(p, d, s + p->prev[i]);
        /* Ty-wide modular compare: :461-464 */
        if (((cur - prv) & mask) != ((uint32_t)p->ed[i] & mask)) {
            int diff = (int)cur - (int)prv;                    /* :467 */
            int sk = tab_get(p, diff);
            if (sk < 1) sk = 1;
            int w = p->wst[i];
            *jump = w < sk ? w : sk;                           /* :536-538 */
            return 0;
        }
    }
    *jump = L - 1 - p->lead;                                   /* :526 */
    return 1;
}

/* The loop of MonkeyMoore<Ty>::search (src/core/monkey_moore.cpp:331-408, :437-544) entered at element `start`
 * and left as soon as the chain reaches element `owned` (or runs out of data): mmo_search is the case
 * start = 0, owned = n.  Slices of ONE chain are what the sliced / multi-GPU search hands from GPU to GPU;
 * *exit_pos receives the position the chain left with.  Positions are relative to data[0]. */
/* Opt-in complete-match mode of the product (mmg_set_complete_matches): every window is evaluated, i.e. the chain
 * advances by 1 whatever F(s) says.  NOT reference behaviour; off by default. */
static int g_complete = 0;
int mmo_set_complete(int on) { int old = g_complete; g_complete = on != 0; return old; }

int64_t mmo_search_slice(const mmo_pattern *p, const void *data, uint64_t n, uint64_t start, uint64_t owned,
                         uint64_t *exit_pos, uint64_t *out_pos, uint32_t *out_vals, uint64_t cap) {
    const uint8_t *d = (const uint8_t *)data;
    uint64_t L = (uint64_t)p->L;
    int64_t count = 0;
    uint64_t s = start;
    while (s < owned && s + L <= n) {
        int jump;
        if (window(p, d, s, &jump)) {
            if ((uint64_t)count < cap) {
                if (out_pos) out_pos[count] = s;
                if (out_vals) {
                    out_vals[2 * count] = p->first_lit < p->L ? load_elem(p, d, s + p->first_lit) : 0;
                    out_vals[2 * count + 1] = p->opp_idx >= 0 ? load_elem(p, d, s + p->opp_idx) : 0;
                }
            }
            count++;
        }
        s += g_complete ? 1u : (uint64_t)jump;
    }
    if (exit_pos) *exit_pos = s;
    return count;
}

int64_t mmo_search(const mmo_pattern *p, const void *data, uint64_t n,
                   uint64_t *out_pos, uint32_t *out_vals, uint64_t cap) {
    return mmo_search_slice(p, data, n, 0, n, NULL, out_pos, out_vals, cap);
}

int mmo_table_size(const mmo_pattern *p) {
    if (p->mode == 2) return 0;
    if (p->nseq == 0) return 2;
    int n = 0;
    for (int i = 0; i < p->nseq; i++) {
        int dup = 0;
        for (int j = 0; j < i; j++) dup |= p->seq[j] == p->seq[i];
        n += !dup;
    }
    return n;
}

void mmo_table(const mmo_pattern *pc, uint32_t v0, uint32_t v1,
               uint32_t *keys, uint32_t *values) {
    mmo_pattern *p = (mmo_pattern *)pc;   /* seq_index may default-insert */
    uint32_t mask = (uint32_t)p->vmax;
    if (p->mode == 2) return;
    if (p->nseq == 0) {
        /* first literal of the (normalised) keyword */
        uint32_t ref = p->mode == 1 ? p->nk[p->first_lit] : p->kw[0];
        uint32_t dist = v0 - ref;                              /* :381, :477-478 */
        uint32_t A = 'A' + dist, a = 'a' + dist;
        if (p->mode == 1 && p->has_case_change) {              /* :487-512 */
            uint32_t od = v1 - p->kw[p->opp_idx];
            if (p->mostly_lowercase) A = 'A' + od; else a = 'a' + od;
        }
        keys[0] = 'A'; values[0] = A & mask;
        keys[1] = 'a'; values[1] = a & mask;
        return;
    }
    /* :387-391, :515-520 */
    uint32_t refc = p->mode == 1 ? p->kw[p->first_lit] : p->kw[0];
    uint32_t dist = v0 - (uint32_t)seq_index(p, refc);
    int n = 0;
    for (int i = 0; i < p->nseq; i++) {
        uint32_t c = p->seq[i];
        int dup = 0;
        for (int j = 0; j < n; j++) dup |= keys[j] == c;
        if (dup) continue;
        keys[n] = c;
        values[n] = ((uint32_t)seq_index(p, c) + dist) & mask;
        n++;
    }
    /* std::map iteration order: ascending key */
    for (int i = 1; i < n; i++) {
        uint32_t k = keys[i], v = values[i];
        int j = i - 1;
        while (j >= 0 && keys[j] > k) { keys[j + 1] = keys[j]; values[j + 1] = values[j]; j--; }
        keys[j + 1] = k; values[j + 1] = v;
    }
}

uint64_t mmo_num_blocks(uint64_t size, uint32_t block_size) {
    /* static_cast<uint32_t>(ceil(double(size)/block)) -- search_engine.cpp:232-234 */
    if (block_size == 0) return 0;
    return (size + block_size - 1) / block_size;
}

typedef struct { uint64_t off; uint32_t v0, v1; } hit_t;

static int hit_cmp(const void *a, const void *b) {
    uint64_t x = ((const hit_t *)a)->off, y = ((const hit_t *)b)->off;
    return x < y ? -1 : x > y;
}

int64_t mmo_engine(const mmo_pattern *p, const uint8_t *file, uint64_t size,
                   uint32_t block_size, int big_endian, int wrap32,
                   uint64_t *out_off, uint32_t *out_vals, uint64_t cap) {
    uint32_t W = p->bits / 8;
    uint32_t overlap = (uint32_t)(p->L - 1) * W;               /* :227 */
    uint32_t full = block_size + overlap;                      /* :230 */
    uint64_t nblocks = mmo_num_blocks(size, block_size);
    if (wrap32) nblocks = (uint32_t)nblocks;

    hit_t *hits = NULL; uint64_t nh = 0, caph = 0;
    uint8_t *work = (uint8_t *)malloc((size_t)full + 8);
    uint64_t *pos = NULL; uint32_t *vals = NULL; uint64_t capm = 0;

    for (uint64_t i = 0; i < nblocks; i++) {
        uint64_t off = wrap32 ? (uint64_t)(uint32_t)((uint32_t)i * block_size) : i * (uint64_t)block_size; /* :242 */
        uint64_t remaining = size - off;                       /* :244 (unsigned) */
        uint32_t bsz = (uint32_t)(full < remaining ? full : remaining);
        /* ifstream::read past EOF yields a short read; bytes beyond stay 0 --
         * only reachable with wrap32 (>= 4 GiB files). */
        for (uint32_t pad = 0; pad < W; pad++) {               /* :129-133 */
            uint64_t count = bsz / W;                          /* :137 */
            if ((uint64_t)pad + count * W > bsz) count -= 1;   /* :139-141 */
            memset(work, 0, bsz);
            if (off < size) {
                uint64_t avail = size - off < bsz ? size - off : bsz;
                memcpy(work, file + off, avail);
            }
            if (W == 2 && big_endian) {                        /* :143-145 */
                for (uint64_t k = 0; k < count; k++) {
                    uint8_t t = work[pad + 2 * k];
                    work[pad + 2 * k] = work[pad + 2 * k + 1];
                    work[pad + 2 * k + 1] = t;
                }
            }
            int64_t m = mmo_search(p, work + pad, count, NULL, NULL, 0);
            if ((uint64_t)m > capm) {
                capm = (uint64_t)m * 2 + 16;
                pos = (uint64_t *)realloc(pos, capm * sizeof(uint64_t));
                vals = (uint32_t *)realloc(vals, capm * 2 * sizeof(uint32_t));
            }
            if (m > 0) mmo_search(p, work + pad, count, pos, vals, capm);
            for (int64_t k = 0; k < m; k++) {
                if (nh == caph) { caph = caph ? caph * 2 : 1024; hits = (hit_t *)realloc(hits, caph * sizeof(hit_t)); }
                hits[nh].off = off + pos[k] * W + pad;         /* :151-154 */
                hits[nh].v0 = vals[2 * k]; hits[nh].v1 = vals[2 * k + 1];
                nh++;
            }
        }
    }
    if (nh) qsort(hits, nh, sizeof(hit_t), hit_cmp);            /* :193-197 */
    for (uint64_t k = 0; k < nh && k < cap; k++) {
        if (out_off) out_off[k] = hits[k].off;
        if (out_vals) { out_vals[2 * k] = hits[k].v0; out_vals[2 * k + 1] = hits[k].v1; }
    }
    free(hits); free(work); free(pos); free(vals);
    return (int64_t)nh;
}
