/*
 * mm_oracle.h -- CPU restatement of Monkey-Moore's relative-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (monkey-moore_b200/).  Only tests/, __graft_entry__
 * .smoke() and bench.py's cpu_baseline / --impl reference legs may use it,
 * and only as the checker / the CPU baseline.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this restatement
 * against every known-answer vector in the reference's own tests
 * (tests/test_monkey_moore.cpp, tests/test_search_engine.cpp) and
 * tests/test_oracle_vs_ref.py checks it differentially against the
 * reference's own sources compiled in place into oracle/_ref/ (see
 * oracle/Makefile).
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef MM_ORACLE_H
#define MM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMO_OK 0
#define MMO_ERR_SKIP_OOB 1   /* "Skip table index out of bounds" (src/core/monkey_moore.cpp:139,274) */
#define MMO_ERR_EMPTY 2      /* assert(!keyword.empty())         (src/core/monkey_moore.cpp:18,28)   */
#define MMO_ERR_HANG 3       /* reference would loop forever (match jump <= 0) -- documented deviation */
#define MMO_ERR_ARG 4

typedef struct mmo_pattern mmo_pattern;

/* MonkeyMoore<Ty>(keyword, wildcard, char_seq)  -- src/core/monkey_moore.cpp:12-22.
 * bits = 8 or 16 selects Ty.  Returns NULL and sets *err on failure. */
mmo_pattern *mmo_compile_keyword(const uint32_t *keyword, int keyword_len,
                                 uint32_t wildcard,
                                 const uint32_t *char_seq, int char_seq_len,
                                 int bits, int *err);

/* MonkeyMoore<Ty>(reference_values)  -- src/core/monkey_moore.cpp:24-39. */
mmo_pattern *mmo_compile_values(const int16_t *values, int n, int bits, int *err);

void mmo_free(mmo_pattern *p);

int mmo_keyword_len(const mmo_pattern *p);
int mmo_elem_bits(const mmo_pattern *p);      /* 8 or 16 */
/* 0 = simple_relative, 1 = wildcard_relative, 2 = value_scan */
int mmo_mode(const mmo_pattern *p);

/* MonkeyMoore<Ty>::search(data, data_len)  -- src/core/monkey_moore.cpp:41-49.
 * data points at n_elems elements of Ty in HOST byte order (may be unaligned).
 * Writes up to cap match positions (element indices) to out_pos and, per
 * match, two raw element values to out_vals[2*i], out_vals[2*i+1]:
 *   [0] the element under the first literal of the (case-normalised) keyword,
 *   [1] the element under the first opposite-case letter (mixed-case keywords
 *       only, else 0).
 * Returns the total number of matches (may exceed cap). */
int64_t mmo_search(const mmo_pattern *p, const void *data, uint64_t n_elems,
                   uint64_t *out_pos, uint32_t *out_vals, uint64_t cap);

/* Number of entries of the equivalency_map a match produces (0 for value scan,
 * 2 for ASCII, |distinct chars of char_seq| otherwise). */
int mmo_table_size(const mmo_pattern *p);

/* Builds the equivalency_map of one match from the two raw values reported by
 * mmo_search -- src/core/monkey_moore.cpp:374-393 and :472-521.  keys/values
 * receive mmo_table_size() entries in ascending key order (std::map order). */
void mmo_table(const mmo_pattern *p, uint32_t v0, uint32_t v1,
               uint32_t *keys, uint32_t *values);

/* SearchEngine<T>::run block decomposition + per-alignment search + sort
 * -- src/core/search_engine.cpp:104-172, 193-197, 218-253.
 * file/size: the whole file image in memory.  block_size =
 * preferred_search_block_size.  big_endian: config.endianness == Big.
 * wrap32 != 0 reproduces the reference's uint32_t block-offset arithmetic
 * (src/core/search_engine.cpp:241-242); wrap32 == 0 uses 64-bit offsets (the
 * documented deviation for >= 4 GiB inputs).
 * Results (file byte offsets, ascending) to out_off/out_vals as in mmo_search.
 * Returns total number of matches. */
int64_t mmo_engine(const mmo_pattern *p, const uint8_t *file, uint64_t size,
                   uint32_t block_size, int big_endian, int wrap32,
                   uint64_t *out_off, uint32_t *out_vals, uint64_t cap);

/* Number of blocks compute_search_blocks() yields (progress-callback count). */
uint64_t mmo_num_blocks(uint64_t size, uint32_t block_size);

/* mm_oracle_stream.c: the engine (64-bit offsets) over blocks [first_block, first_block + nblocks) of a SYNTHETIC
 * file that is regenerated block by block (counter-based generator of monkey-moore_b200/synth.py + planted
 * patches, ascending by offset), folded into an order-sensitive digest out3 = {count, S0, S1}; optionally the
 * ordered list itself (out_off / out_vals, up to list_cap matches).  nthreads host threads work on disjoint blocks. */
uint64_t mmo_digest_mix(uint64_t off, uint32_t v0, uint32_t v1);
int64_t mmo_engine_synth(const mmo_pattern *p, uint64_t seed, uint32_t byte_mask, uint64_t total_size,
                         uint32_t block_size, uint64_t first_block, uint64_t nblocks, int big_endian,
                         const uint64_t *patch_off, const uint32_t *patch_len, const uint8_t *patch_bytes,
                         uint64_t npatches, int nthreads, uint64_t *out3,
                         uint64_t *out_off, uint32_t *out_vals, uint64_t list_cap);

#ifdef __cplusplus
}
#endif
#endif
