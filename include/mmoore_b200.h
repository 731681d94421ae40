/*
 * mmoore_b200.h -- C-ABI of the B200-native relative-search path.
 *
 * This is the drop-in boundary underneath the reference's link-time C++ API
 * (include/mmoore/{monkey_moore,search_engine}.hpp, static library monkey-core,
 * /root/reference/src/core/CMakeLists.txt:1-3).  The reference has no FFI of its
 * own; each entry point below names the reference interface it stands in for
 * (file:line under /root/reference/).  The C++ classes in include/mmoore/ are thin
 * wrappers over these calls; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; integer status
 * codes (0 == MMG_OK), never exceptions; the caller owns every input buffer;
 * the library owns a results object until mmg_results_free().  Handles are
 * immutable after creation and may be shared between threads; every scan call
 * runs on its own CUDA stream.  There is NO CPU fallback: every scan entry
 * point fails with MMG_ERR_CUDA when no CUDA device is usable.
 */
#ifndef MMOORE_B200_H
#define MMOORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMG_OK 0
#define MMG_ERR_SKIP_OOB 1  /* -> std::runtime_error("Skip table index out of bounds"), src/core/monkey_moore.cpp:139,274 */
#define MMG_ERR_EMPTY 2     /* empty keyword / value list (assert at src/core/monkey_moore.cpp:18,28) */
#define MMG_ERR_HANG 3      /* pattern whose match advance is <= 0: the reference never terminates (:398,:526); rejected */
#define MMG_ERR_ARG 4       /* bad argument */
#define MMG_ERR_CUDA 5      /* CUDA runtime failure or no device (message in mmg_last_error) */
#define MMG_ERR_TOO_LONG 6  /* keyword longer than 32767 elements (more than 128: accepted, per-chain kernels only) */
#define MMG_ERR_NOMEM 7

#define MMG_MAX_KEYWORD 32767          /* elements; keywords of more than MMG_FAST_KEYWORD run on the per-chain kernels */
#define MMG_FAST_KEYWORD 128

typedef struct mmg_program mmg_program;   /* a compiled pattern (== one MonkeyMoore<Ty> instance) */
typedef struct mmg_results mmg_results;   /* the match list of one scan */

/* Human-readable description of the last error on this thread. */
const char *mmg_last_error(void);

/* Number of usable CUDA devices (0 => every scan call fails with MMG_ERR_CUDA). */
int mmg_device_count(void);

/* MonkeyMoore<Ty>::MonkeyMoore(keyword, wildcard, char_seq)
 *   include/mmoore/monkey_moore.hpp:29-33, src/core/monkey_moore.cpp:12-22, 54-304.
 * keyword/char_seq are UTF-32 code points (CharType == char32_t); elem_bits is 8 or
 * 16 (the two instantiations, src/core/monkey_moore.cpp:587-588). */
int mmg_program_create_keyword(const uint32_t *keyword, int keyword_len, uint32_t wildcard,
                               const uint32_t *char_seq, int char_seq_len, int elem_bits,
                               mmg_program **out);

/* MonkeyMoore<Ty>::MonkeyMoore(const std::vector<short>& reference_values)  (value scan)
 *   include/mmoore/monkey_moore.hpp:40-42, src/core/monkey_moore.cpp:24-39. */
int mmg_program_create_values(const int16_t *values, int n, int elem_bits, mmg_program **out);

void mmg_program_free(mmg_program *p);

/* Introspection used by the host wrappers and the tests. */
int mmg_program_keyword_len(const mmg_program *p);
int mmg_program_mode(const mmg_program *p);          /* 0 simple_relative, 1 wildcard_relative, 2 value_scan */
int mmg_program_table_size(const mmg_program *p);    /* entries of the equivalency_map of a match */

/* equivalency_map of one match (src/core/monkey_moore.cpp:374-393, 472-521) from the two raw
 * element values a scan reports for it.  keys/values receive mmg_program_table_size()
 * entries in ascending key order (std::map iteration order). */
void mmg_program_table(const mmg_program *p, uint32_t v0, uint32_t v1, uint32_t *keys, uint32_t *values);

/* Where the bytes live. */
#define MMG_MEM_HOST 0     /* host pointer: the call copies host->device (and is what `e2e` times) */
#define MMG_MEM_DEVICE 1   /* device pointer on the current device, 16-byte aligned */

/* MonkeyMoore<Ty>::search(const Ty* data, uint64_t data_len)
 *   include/mmoore/monkey_moore.hpp:51, src/core/monkey_moore.cpp:41-49, 316-546.
 * data_len is in ELEMENTS, elements are in host byte order; one chain from element 0.
 * Match positions are element indices, ascending. */
int mmg_search(const mmg_program *p, const void *data, uint64_t data_len, int mem, mmg_results **out);

/* MonkeyMoore<Ty>::search over a buffer that is cut into SLICES (one per GPU, or per call when the buffer
 * does not fit): still ONE chain from element 0 of the whole buffer (src/core/monkey_moore.cpp:316-546), so
 * where the chain enters a slice depends on everything before it.  Every slice is scanned without that
 * knowledge; what it hands on is its MAP: for each possible entry phase (distance of the first chain position
 * from the slice's first element, < mmg_program_max_jump) the phase with which the chain enters the next slice.
 *
 *   mmg_chain_begin   enqueues filter + maps of one slice.  data[0] is element first_element of the buffer; the
 *                     slice OWNS the windows that start in its first owned_len elements; avail_len >= owned_len
 *                     elements are present: owned_len + keyword_len - 1 (or more) when another slice follows --
 *                     owned_len * sizeof(Ty) must then be a multiple of 4096 -- and exactly owned_len for the
 *                     last slice.  *out is a results object that is not yet usable as one.
 *   mmg_chain_map     waits for the map of the slice: map[e] = exit phase for entry phase e, *n entries.
 *   mmg_chain_entry   host arithmetic: the entry phase of slice `nslices` given the maps of slices 0..nslices-1
 *                     (maps + k * stride); 0 for the first slice.
 *   mmg_chain_finish  enqueues the rest of the scan with the slice's entry phase; from then on *out behaves like
 *                     the results of mmg_engine_scan_async (element indices of the WHOLE buffer, ascending).
 * The concatenation of the slices' lists in slice order equals mmg_search over the whole buffer.
 * mmg_comm_search (below) runs this protocol across the ranks of a communicator. */
int mmg_chain_begin(const mmg_program *p, const void *data, uint64_t owned_len, uint64_t avail_len, int mem,
                    uint64_t first_element, mmg_results **out);
int mmg_chain_map(mmg_results *r, uint8_t *map, int capacity, int *n);
uint32_t mmg_chain_entry(const uint8_t *maps, int stride, int nslices);
int mmg_chain_finish(mmg_results *r, uint32_t entry_phase);
int mmg_program_max_jump(const mmg_program *p);      /* number of entry phases of a slice (<= 128) */

/* The chunk engine of mmoore::SearchEngine<T>::run WITHOUT file I/O, previews and callbacks:
 *   compute_search_blocks          src/core/search_engine.cpp:218-253
 *   per-block, per-alignment scan  src/core/search_engine.cpp:129-159
 *   merge + sort by offset         src/core/search_engine.cpp:193-197
 * bytes/nbytes   : the file image, or the contiguous slice of it that starts at file
 *                  offset first_block*block_size (multi-GPU sharding, one slice per rank).
 * file_size      : size of the WHOLE file (needed to shape the last block).
 * block_size     : SearchConfig::preferred_search_block_size.
 * first_block,
 * num_blocks     : which blocks of the file this call scans; the slice must cover
 *                  [first_block*block_size, min(file_size, (first_block+num_blocks)*block_size + overlap)).
 *                  num_blocks == 0 means "all blocks from first_block to the end".
 * big_endian     : SearchConfig::endianness == Endianness::Big (16-bit only).
 * Offsets reported are FILE offsets (64-bit; the reference's uint32_t wrap past 4 GiB,
 * src/core/search_engine.cpp:241-242, is deliberately not reproduced), ascending. */
int mmg_engine_scan(const mmg_program *p, const void *bytes, uint64_t nbytes, int mem,
                    uint64_t file_size, uint32_t block_size, uint64_t first_block, uint64_t num_blocks,
                    int big_endian, mmg_results **out);

/* Asynchronous form: enqueues the scan on the calling thread's stream and returns at once with a
 * PENDING results object, so that the host work of the next scan overlaps the GPU work of this one.
 * mmg_results_wait() (or any accessor, or mmg_results_free) completes it and returns its status.  The
 * input bytes must stay valid until then. */
int mmg_engine_scan_async(const mmg_program *p, const void *bytes, uint64_t nbytes, int mem,
                          uint64_t file_size, uint32_t block_size, uint64_t first_block, uint64_t num_blocks,
                          int big_endian, mmg_results **out);
int mmg_results_wait(mmg_results *r);

/* Number of blocks compute_search_blocks() produces (== number of per-block progress callbacks,
 * tests/test_search_engine.cpp:376-396). */
uint64_t mmg_num_blocks(uint64_t file_size, uint32_t block_size);

/* Results: count, then copies into caller buffers (offsets: count u64; values: 2*count u32,
 * [2i] = element under the first literal, [2i+1] = element under the first opposite-case letter). */
uint64_t mmg_results_count(const mmg_results *r);
int mmg_results_copy(const mmg_results *r, uint64_t first, uint64_t n, uint64_t *offsets, uint32_t *values);
/* Device-resident views (valid until mmg_results_free): offsets as u64[count], values packed
 * as u32[count] = v0 | v1 << 16.  For NCCL gathers without a host round trip. */
const uint64_t *mmg_results_device_offsets(const mmg_results *r);
const uint32_t *mmg_results_device_values(const mmg_results *r);
/* Distinct inferred tables of a match list -- the GUI's default "one row per table" view
 * (src/gui/monkey_frame.cpp:1236-1245: a result is listed iff no earlier result has an equal values map).
 * Writes to `indices` (room for `capacity` entries; may be NULL) the ascending match-list indices of the FIRST match
 * of every distinct table and stores their number in *n_unique (call with capacity 0 to size the buffer).  The
 * reduction runs on the device over the values the scan emitted; the match list does not travel to the host. */
int mmg_results_unique(const mmg_program *p, const mmg_results *r, uint64_t *indices, uint64_t capacity, uint64_t *n_unique);
void mmg_results_free(mmg_results *r);

/* Timing / accounting of the last scan that produced `r` (for bench.py). */
typedef struct mmg_scan_stats {
    float ms_total;          /* device time of the whole scan (CUDA events on the scan's stream), excl. H2D */
    float ms_filter;         /* the streaming filter kernel (the HBM-bound one) */
    float ms_h2d;            /* host->device copy, 0 for MMG_MEM_DEVICE */
    uint32_t launches;       /* kernels launched by this scan */
    uint32_t fast_path;      /* 1 = tiled streaming path, 0 = generic per-chain path */
    uint64_t events;         /* candidate windows that needed exact evaluation */
    uint64_t bytes_scanned;
    uint32_t resolve_kind;   /* 0 resolve kernel, 1 resolved inside the filter kernel (sparse scans), 2 that was tried and the resolve kernel re-ran */
    uint32_t chain_entry;    /* chain slices (mmg_chain_*, mmg_comm_search): the entry phase the slice was finished with */
} mmg_scan_stats;
int mmg_results_stats(const mmg_results *r, mmg_scan_stats *out);

/* ---- multi-GPU: gather of the match lists to rank 0 over NCCL (one process per GPU) -----------------
 * The reference's engine merges per-block results of its thread pool and sorts them
 * (src/core/search_engine.cpp:82-102, 193-197); across GPUs ranks scan disjoint block ranges
 * (mmg_engine_scan with first_block/num_blocks) and only the result lists travel.  Rank order equals
 * file order, so the concatenation is already sorted.  NCCL is loaded lazily (dlopen): single-GPU
 * users do not need it.  All gather work runs on the communicator's own CUDA stream.
 *   mmg_comm_unique_id : rank 0 creates the 128-byte NCCL id; the caller broadcasts it (any transport).
 *   mmg_comm_create    : collective over all ranks; capacity = entries of the fixed packed buffer.
 *   mmg_comm_gather    : collective; the `nlists` lists of one step (same nlists on every rank) go to
 *                        rank 0 in ONE grouped NCCL operation.  Whatever exceeds the packed capacity is set
 *                        aside and travels at the start of every rank's NEXT call into the communicator
 *                        (mmg_comm_gather or mmg_comm_wait), so no send ever waits for a peer outside a
 *                        matching collective call.  The call only enqueues work.  *out is non-NULL on rank 0
 *                        only; using it (count / copy / pieces / free) completes it, which for lists that
 *                        did not fit needs the other ranks' next call -- call mmg_comm_wait on every rank
 *                        first when in doubt.  The lists may be freed right after the call on every rank.
 *   mmg_comm_wait      : collective; sends / receives what the last gather set aside and blocks until this
 *                        rank's part of every gather so far has executed; *ms_last
 *                        (may be NULL) = device time of the last gather on this rank's gather stream.
 *   mmg_gathered_pieces: device-resident pieces (rank order == file order) of one gathered list. */
typedef struct mmg_comm mmg_comm;
typedef struct mmg_gathered mmg_gathered;
int mmg_comm_unique_id(void *out128);
int mmg_comm_create(const void *id128, int rank, int world, uint64_t capacity, mmg_comm **out);
void mmg_comm_destroy(mmg_comm *c);
int mmg_comm_gather(mmg_comm *c, const mmg_results *const *lists, int nlists, mmg_gathered **out);
int mmg_comm_wait(mmg_comm *c, float *ms_last);
/* One chain over a buffer that is spread over the ranks (rank r holds slice r, see mmg_chain_begin): every rank
 * scans its slice, the slice maps are exchanged with one all-gather of max_jump bytes per rank -- the only step
 * of the whole path where ranks depend on each other -- and every rank finishes its slice with the entry phase
 * composed from the maps of the ranks before it.  *out holds this rank's part of the result of
 * MonkeyMoore<Ty>::search on the whole buffer; mmg_comm_gather concatenates the parts on rank 0.  Collective. */
int mmg_comm_search(mmg_comm *c, const mmg_program *p, const void *data, uint64_t owned_len, uint64_t avail_len,
                    int mem, uint64_t first_element, mmg_results **out);
uint64_t mmg_gathered_count(const mmg_gathered *g, int list);
int mmg_gathered_copy(const mmg_gathered *g, int list, uint64_t *offsets, uint32_t *values);
int mmg_gathered_pieces(const mmg_gathered *g, int list, const uint64_t **offs, const uint32_t **vals, uint64_t *ns, int cap);
void mmg_gathered_free(mmg_gathered *g);

/* Page-locked host staging memory for file -> HBM ingestion (SearchEngine<T>::run reads the file into
 * such a buffer so that the H2D copy runs at PCIe speed).  NULL when allocation fails / no device. */
void *mmg_host_alloc(uint64_t nbytes);
void mmg_host_free(void *p);
/* memcpy split over the library's pool of host threads (file pages / pageable memory -> staging memory at more than one
 * core's copy bandwidth).  Host pointers passed to the scan calls with MMG_MEM_HOST need no preparation: page-locked
 * ones are copied directly, pageable ones of 4 MiB and more go through the library's own pinned staging ring. */
int mmg_host_copy(void *dst, const void *src, uint64_t nbytes);

/* Stream selection for the calling thread: use_it != 0 makes every later scan of this thread run on
 * `cuda_stream` (a cudaStream_t; NULL is the legacy default stream), so a caller can bracket scans with
 * its own CUDA events; use_it == 0 returns to the library's private non-blocking stream. */
int mmg_set_stream(void *cuda_stream, int use_it);

/* Synthetic ROM generator used by bench.py and the tests (no reference counterpart; SURVEY.md
 * section 8d): byte i of the stream = byte (i mod 8) of splitmix64(seed ^ (i / 8)), AND byte_mask.
 * Fills device memory [device_ptr, device_ptr + nbytes) with stream bytes first_byte ...;
 * nbytes, first_byte and the pointer must be multiples of 8. */
int mmg_synth_fill(void *device_ptr, uint64_t nbytes, uint64_t seed, uint64_t first_byte, uint32_t byte_mask);

/* Testing knob (returns the previous mode): 0 = automatic; 1 = force the per-chain generic kernels;
 * 2 = tiled path with exact evaluation of every window (no SWAR filter); 4 = tiled path, never resolve inside the
 * filter kernel (sparse scans then run the separate resolve kernel like dense ones). */
int mmg_set_path_override(int mode);

/* Opt-in SUPERSET of the reference's result (not the parity target): with on != 0 every later scan of
 * the process reports EVERY window that matches -- per block and alignment, ascending -- instead of only the matches
 * the reference's skip chain happens to visit (src/core/monkey_moore.cpp:398-404, :526-541: after a miss the chain
 * advances by a table value and may jump over a true match).  The filter evaluates all windows anyway; only the
 * replay of the chain is skipped.  Returns the previous setting.  Default: off (bit-exact with the reference). */
int mmg_set_complete_matches(int on);

#ifdef __cplusplus
}
#endif
#endif /* MMOORE_B200_H */
