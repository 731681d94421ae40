// scan_kernels.cuh -- device-side data model of the B200 relative-search path.
//
// The reference walks each (block, alignment) buffer with a sequential, lossy Boyer-Moore
// chain  s <- s + jump(s)  (/root/reference/src/core/monkey_moore.cpp:347-405, 449-541) and
// restarts it at every engine block (/root/reference/src/core/search_engine.cpp:129-159).
// Bit-exact results therefore need that chain replayed.  The GPU formulation:
//
//   K1  filter      streams the bytes once (16-byte loads), computes the element-difference
//                   stream SWAR-style and flags the few windows whose first comparison does
//                   not end in the default advance J0.  Flagged windows are evaluated exactly
//                   and become EVENTS (window start, advance, match bit), grouped per sub-tile.
//   K2  resolve     one warp per engine block: per sub-tile and alignment class the map  entry phase ->
//                   exit phase  of the chain through the sub-tile's events (everything else advances by
//                   J0); composition of the maps along the block's chains; replay of the TRUE chains
//                   through the events (marks the matches they visit); decoupled look-back over the
//                   per-block match counts; ordered emission of file offsets + table base values.
//
// Irregular geometries (block size not a multiple of the sub-tile) use the per-chain kernels G*.
#ifndef MMG_SCAN_KERNELS_CUH
#define MMG_SCAN_KERNELS_CUH

#include "program.h"

#include <cuda_runtime.h>
#include <stdint.h>

#define MMG_SUBTILE 4096u          // bytes of window starts per sub-tile
#define MMG_SUBTILE_SHIFT 12
#define MMG_ROW 512u               // bytes one warp loads per step (32 lanes x 16 B)
#define MMG_ROWS_PER_SUB (MMG_SUBTILE / MMG_ROW)

// event word: [11:0] window start (byte offset in the sub-tile)  [23:16] advance  [24] match  [25] visited
#define MMG_EV_OFF(e) ((e) & 0xFFFu)
#define MMG_EV_JUMP(e) (((e) >> 16) & 0xFFu)
#define MMG_EV_MATCH 0x01000000u
#define MMG_EV_VISITED 0x02000000u

struct MmgGeom {
    const uint8_t *data;   // device bytes of the slice; data[0] is file offset base_offset
    uint64_t S;            // valid bytes in the slice
    uint64_t B;            // block size (bytes).  search(): one block covering everything
    uint64_t base_offset;  // added to every reported offset (bytes); search() divides by W afterwards
    uint32_t nblocks;
    uint32_t ov;           // (L-1)*W overlap bytes
    uint32_t npads;        // alignments searched per block: W for the engine, 1 for search()
    uint32_t big_endian;
    uint32_t report_shift; // 0: report byte offsets; 1: report element indices of a 16-bit search()
    uint32_t spb;          // sub-tiles per block (fast path)
    uint32_t nsub;         // total sub-tiles (fast path)
    uint32_t chunk_subs;   // sub-tiles per warp work unit (divides spb)
    uint32_t nchunks;
    uint32_t l2_hint;      // L2 policy of the input stream: 0 none, 1 evict_first, 2 evict_unchanged (tma_load_1d)
    uint32_t static_chunks; // != 0: warp w takes chunks w, w + warps, ... instead of drawing them (small inputs)
    // resolve geometry: one CTA per SEGMENT of at most 128 sub-tiles; a block of more than 128 sub-tiles (search() on a
    // large buffer, the GUI's 8 MiB blocks) is cut into segs_per_block segments whose entry phases come from a prefix
    // over the segment maps (k_resolve<MAPS_ONLY>, then k_segphase or -- one block of many segments -- k_rangemap + k_chainphase)
    uint32_t segs_per_block;
    uint32_t nseg;         // total segments = CTAs of the resolve kernel
    // slice of a longer chain (mmg_chain_*): the one block of the scan is entered with these phases (per alignment class)
    // instead of 0, and every segment -- the first one too -- reads its entry phase from segphase (k_chainphase)
    uint32_t chain;
    uint32_t entry[2];
    // opt-in superset of the reference's result (mmg_set_complete_matches): EVERY window that matches is
    // reported, not only those the lossy skip chain happens to visit
    uint32_t complete;
};

struct MmgScratch {
    uint32_t *ev;          // event words, one private region per filter warp
    uint32_t ev_per_warp;
    uint2 *ext;            // [nsub] {first event, number of events} of the sub-tile, written for EVERY sub-tile by the filter
    uint32_t *mcount;      // [nsub] visited matches (valid where the sub-tile has events)
    uint64_t *mbase;       // [nsub] position of the sub-tile's first match in the output (ditto)
    // ---- fused resolve of sparse scans (scan_kernels.cu, "Resolve FUSED into the filter kernels")
    uint32_t fuse;         // != 0: the filter kernel resolves the engine blocks itself
    uint32_t *brec;        // [nblocks][32] per-block record {match count, byte offsets of the matches}
    uint32_t *bcount;      // [nblocks] match count of the block (compact copy for the final prefix)
    uint64_t ev_total;     // entries of `ev`
    uint64_t *out_off;     // the scan's result buffers (the last warp of the grid writes the matches)
    uint32_t *out_val;
    uint64_t capacity;
    // Zero state: all zero when a scan starts.  The last CTA of the resolve kernel to finish copies `status` to the
    // host's pinned slot and zeroes it all again, so a workspace serves scan after scan without a memset.
    uint64_t *status;      // [0] events needed by the fullest warp region (overflow check) [1] total events
                           // [2] total matches [3] next chunk (dynamic scheduling)
    uint64_t *lookback;    // [nblocks] decoupled look-back words of the per-block match counts
    uint32_t *ticket;      // [0] block ticket of the resolve kernel  [1] CTAs of the resolve kernel that are done
                           // [2] fused resolve: a block held more events / matches than it can stage
                           // fused resolve: [1], [3] grid barriers (CTAs that left the filter loop / resolved their blocks),
                           // [4] CTAs that are done
    uint64_t *host_status; // pinned, device-visible: receives status[0..3] (+ [4] = ticket[2]) when the resolve kernel ends
    uint8_t *segmap;       // [nseg][2][jp] entry phase -> exit phase of a whole segment (only when segs_per_block > 1)
    uint8_t *segphase;     // [nseg][2] entry phase of the segment
    uint8_t *rangemap;     // [<= 128][2][jp] chain slices: composed maps of ranges of segments
    uint8_t *slicemap_host; // pinned [2][jp]: the map of the whole slice
};

#endif
