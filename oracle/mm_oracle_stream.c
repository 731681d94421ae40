/*
 * mm_oracle_stream.c -- the oracle's engine applied to a SYNTHETIC file that is never materialised.
 *
 * TEST INFRASTRUCTURE ONLY (see mm_oracle.h).  bench.py and the GPU tests use it to check full-size runs
 * (4 GiB on one GPU, 16 / 64 GiB over several) whose blobs exist only in HBM: the bytes of every engine
 * block are regenerated on the host with the same counter-based generator the device uses
 * (monkey-moore_b200/synth.py: byte i = byte (i mod 8) of splitmix64(seed ^ (i / 8)), AND byte_mask;
 * planted patches on top), each block is searched exactly like mmo_engine() does
 * (/root/reference/src/core/search_engine.cpp:129-159 per block and alignment, 64-bit block offsets,
 * i.e. wrap32 == 0), and the ordered match list is folded into an ORDER-SENSITIVE digest
 *
 *     h_i = mix64(offset_i + 0x9E3779B97F4A7C15 * (value_word_i + 1)),  value_word = v0 | v1 << 16
 *     S0  = sum h_i          S1 = sum h_i * (2 i + 1)        (mod 2^64, i = index in this range's list)
 *
 * Lists of adjacent block ranges compose:  S1(A ++ B) = S1(A) + S1(B) + 2 |A| S0(B), so every rank can check
 * its own shard and rank 0 can check the gathered whole.  Blocks are independent chains, so they are
 * processed by a small pthread pool.
 */
#include "mm_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

uint64_t mmo_digest_mix(uint64_t off, uint32_t v0, uint32_t v1) {
    const uint64_t word = (uint64_t)((v0 & 0xFFFFu) | (v1 << 16));
    return splitmix64(off + 0x9E3779B97F4A7C15ull * (word + 1));
}

typedef struct {
    const mmo_pattern *p;
    uint64_t seed, total_size, first_block, nblocks;
    uint32_t byte_mask, block_size;
    int big_endian;
    const uint64_t *patch_off;      /* ascending */
    const uint32_t *patch_len;
    const uint64_t *patch_at;       /* start of patch k's bytes in patch_bytes */
    const uint8_t *patch_bytes;
    uint64_t npatches;
    /* per block results */
    uint64_t *cnt, *s0, *s1;
    /* optional full list (offsets, 2 values per match): filled when list_cap allows */
    uint64_t **blk_off; uint32_t **blk_val; int keep_lists;
    volatile uint64_t next;         /* work counter */
    pthread_mutex_t mu;
} job_t;

static void gen_bytes(const job_t *j, uint64_t first, uint64_t n, uint8_t *out) {
    /* bytes [first, first + n) of the synthetic file */
    uint64_t i = 0;
    const uint64_t mask8 = 0x0101010101010101ull * (uint64_t)(j->byte_mask & 0xFFu);
    while (i < n) {
        const uint64_t b = first + i, w = b >> 3, r = b & 7;
        uint64_t v = splitmix64(j->seed ^ w) & mask8;
        if (r == 0 && n - i >= 8) {
            memcpy(out + i, &v, 8);             /* little-endian hosts only (x86-64 / aarch64) */
            i += 8;
        } else {
            out[i] = (uint8_t)(v >> (8 * r));
            i += 1;
        }
    }
    /* planted patches that overlap [first, first + n) */
    uint64_t lo = 0, hi = j->npatches;
    while (lo < hi) {            /* first patch whose end may reach `first`: patches are short, scan back a little */
        const uint64_t mid = (lo + hi) >> 1;
        if (j->patch_off[mid] + j->patch_len[mid] <= first) lo = mid + 1; else hi = mid;
    }
    /* patch ends are not monotone in general; back up while earlier patches still overlap */
    while (lo > 0 && j->patch_off[lo - 1] + j->patch_len[lo - 1] > first) lo--;
    for (uint64_t k = lo; k < j->npatches && j->patch_off[k] < first + n; k++) {
        const uint64_t po = j->patch_off[k], pl = j->patch_len[k];
        const uint64_t a = po > first ? po : first, b = po + pl < first + n ? po + pl : first + n;
        if (a < b) memcpy(out + (a - first), j->patch_bytes + j->patch_at[k] + (a - po), b - a);
    }
}

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    const uint32_t bits = (uint32_t)mmo_elem_bits(j->p), EW = bits / 8;
    const uint32_t overlap = (uint32_t)(mmo_keyword_len(j->p) - 1) * EW;
    const uint64_t full = (uint64_t)j->block_size + overlap;
    uint8_t *raw = (uint8_t *)malloc(full + 16), *work = (uint8_t *)malloc(full + 16);
    const uint64_t cap = full / EW + 8;
    uint64_t *pos[2]; uint32_t *val[2];
    for (int k = 0; k < 2; k++) { pos[k] = (uint64_t *)malloc(cap * sizeof(uint64_t)); val[k] = (uint32_t *)malloc(cap * 2 * sizeof(uint32_t)); }
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const uint64_t bi = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (bi >= j->nblocks) break;
        const uint64_t off = (j->first_block + bi) * (uint64_t)j->block_size;
        const uint64_t remaining = j->total_size - off;
        const uint64_t bsz = full < remaining ? full : remaining;
        gen_bytes(j, off, bsz, raw);
        int64_t m[2] = {0, 0};
        for (uint32_t pad = 0; pad < EW; pad++) {                  /* search_engine.cpp:129-159 */
            uint64_t count = bsz / EW;
            if ((uint64_t)pad + count * EW > bsz) count -= 1;
            const uint8_t *src = raw;
            if (EW == 2 && j->big_endian) {
                memcpy(work, raw, bsz);
                for (uint64_t k = 0; k < count; k++) {
                    const uint8_t t = work[pad + 2 * k];
                    work[pad + 2 * k] = work[pad + 2 * k + 1];
                    work[pad + 2 * k + 1] = t;
                }
                src = work;
            }
            m[pad] = mmo_search(j->p, src + pad, count, pos[pad], val[pad], cap);
            for (int64_t k = 0; k < m[pad]; k++) pos[pad][k] = off + pos[pad][k] * EW + pad;   /* :151-154 */
        }
        /* merge the (at most two) ascending lists by offset == the engine's final sort restricted to this block */
        const uint64_t n = (uint64_t)(m[0] + m[1]);
        uint64_t s0 = 0, s1 = 0, a = 0, b = 0;
        uint64_t *lo_ = NULL; uint32_t *lv_ = NULL;
        if (j->keep_lists && n) { lo_ = (uint64_t *)malloc(n * sizeof(uint64_t)); lv_ = (uint32_t *)malloc(n * 2 * sizeof(uint32_t)); }
        for (uint64_t i = 0; i < n; i++) {
            int take0 = b >= (uint64_t)m[1] || (a < (uint64_t)m[0] && pos[0][a] < pos[1][b]);
            const uint64_t o = take0 ? pos[0][a] : pos[1][b];
            const uint32_t v0 = take0 ? val[0][2 * a] : val[1][2 * b], v1 = take0 ? val[0][2 * a + 1] : val[1][2 * b + 1];
            if (take0) a++; else b++;
            const uint64_t h = mmo_digest_mix(o, v0, v1);
            s0 += h;
            s1 += h * (2 * i + 1);
            if (lo_) { lo_[i] = o; lv_[2 * i] = v0; lv_[2 * i + 1] = v1; }
        }
        j->cnt[bi] = n; j->s0[bi] = s0; j->s1[bi] = s1;
        if (j->keep_lists) { j->blk_off[bi] = lo_; j->blk_val[bi] = lv_; }
    }
    for (int k = 0; k < 2; k++) { free(pos[k]); free(val[k]); }
    free(raw); free(work);
    return NULL;
}

/* Engine over blocks [first_block, first_block + nblocks) of the synthetic file.  out3 = {count, S0, S1}.
 * When out_off != NULL the ordered list itself is written as well (up to list_cap matches).
 * Returns the number of matches, or -1 on bad arguments. */
int64_t mmo_engine_synth(const mmo_pattern *p, uint64_t seed, uint32_t byte_mask, uint64_t total_size,
                         uint32_t block_size, uint64_t first_block, uint64_t nblocks, int big_endian,
                         const uint64_t *patch_off, const uint32_t *patch_len, const uint8_t *patch_bytes,
                         uint64_t npatches, int nthreads, uint64_t *out3,
                         uint64_t *out_off, uint32_t *out_vals, uint64_t list_cap) {
    if (!p || block_size == 0 || !out3) return -1;
    const uint64_t all = mmo_num_blocks(total_size, block_size);
    if (first_block > all) return -1;
    if (nblocks == 0 || first_block + nblocks > all) nblocks = all - first_block;
    job_t j;
    memset(&j, 0, sizeof(j));
    j.p = p; j.seed = seed; j.byte_mask = byte_mask; j.total_size = total_size; j.block_size = block_size;
    j.first_block = first_block; j.nblocks = nblocks; j.big_endian = big_endian;
    j.patch_off = patch_off; j.patch_len = patch_len; j.patch_bytes = patch_bytes; j.npatches = npatches;
    uint64_t *at = (uint64_t *)malloc((npatches + 1) * sizeof(uint64_t));
    at[0] = 0;
    for (uint64_t k = 0; k < npatches; k++) at[k + 1] = at[k] + patch_len[k];
    j.patch_at = at;
    j.cnt = (uint64_t *)calloc(nblocks + 1, sizeof(uint64_t));
    j.s0 = (uint64_t *)calloc(nblocks + 1, sizeof(uint64_t));
    j.s1 = (uint64_t *)calloc(nblocks + 1, sizeof(uint64_t));
    j.keep_lists = out_off != NULL;
    if (j.keep_lists) {
        j.blk_off = (uint64_t **)calloc(nblocks + 1, sizeof(uint64_t *));
        j.blk_val = (uint32_t **)calloc(nblocks + 1, sizeof(uint32_t *));
    }
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > nblocks && nblocks > 0) nthreads = (int)nblocks;
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    uint64_t n = 0, s0 = 0, s1 = 0;
    for (uint64_t b = 0; b < nblocks; b++) {
        s1 += j.s1[b] + 2 * n * j.s0[b];
        s0 += j.s0[b];
        if (j.keep_lists && j.cnt[b]) {
            for (uint64_t i = 0; i < j.cnt[b] && n + i < list_cap; i++) {
                out_off[n + i] = j.blk_off[b][i];
                if (out_vals) { out_vals[2 * (n + i)] = j.blk_val[b][2 * i]; out_vals[2 * (n + i) + 1] = j.blk_val[b][2 * i + 1]; }
            }
            free(j.blk_off[b]); free(j.blk_val[b]);
        }
        n += j.cnt[b];
    }
    out3[0] = n; out3[1] = s0; out3[2] = s1;
    free(th); free(at); free(j.cnt); free(j.s0); free(j.s1); free(j.blk_off); free(j.blk_val);
    pthread_mutex_destroy(&j.mu);
    return (int64_t)n;
}
