// program.h -- the compiled pattern ("pattern program") shared by host and device.
//
// A MonkeyMoore<Ty> instance of the reference (include/mmoore/monkey_moore.hpp:18-94 under
// /root/reference/) owns a dense skip table of 2*(max+1) ints (512 KiB for 16-bit,
// src/core/monkey_moore.cpp:63-64) plus per-position vectors.  On the GPU the same
// information travels as one small POD that fits in kernel-parameter (constant) space:
// the ordered list of comparisons a window performs, a SPARSE skip table (only <= L
// entries differ from the default) and the handful of scalars the chain replay needs.
#ifndef MMG_PROGRAM_H
#define MMG_PROGRAM_H

#include <stdint.h>

#define MMG_MAXL 128

// One comparison of a window, in evaluation order (right to left, literals only; the first
// literal's wrap-around comparison can never be the first to fail -- the differences
// telescope -- so it is not listed).
struct MmgCheck {
    int16_t i;     // keyword index of the "current" element
    int16_t lag;   // i - prev(i) > 0: distance (elements) to the element it is compared against
    int32_t ed;    // expected signed difference
    int32_t cap;   // wildcard_skip_table[i] (wildcard mode) or INT32_MAX
};

struct MmgProgram {
    int32_t W;            // bytes per element (1 or 2)
    int32_t L;            // keyword length
    int32_t modular;      // 1: differences compared modulo 2^(8W) (wildcard mode, src/core/monkey_moore.cpp:461-464)
                          // 0: exact signed differences (simple / value scan, :336-345)
    int32_t ncheck;
    int32_t ntab;
    int32_t tab_default;  // skip for a difference that is not a table key, already max(.,1)
    int32_t match_jump;   // advance after a match: L-1-leading_wildcards
    int32_t J0;           // advance when the FIRST comparison fails on a difference outside `keys`
    int32_t Jmax;         // largest advance any window can produce
    int32_t first_lit;    // keyword index whose element yields the table base value (v0)
    int32_t opp_idx;      // keyword index of the first opposite-case letter (v1), or -1
    int32_t nkeys;        // filter keys; -1 => every window must be evaluated exactly
    // 16-bit pre-filter, range stage (multi-key patterns whose keys cluster, e.g. differences of character-sequence
    // indices): every key k satisfies (k - rng_lo) mod 2^16 <= rng_w.  rng_c = the SWAR constant (upper half 1 - rng_lo,
    // lower half -rng_lo); rng_w = 0xFFFFFFFF disables the stage.
    uint32_t rng_c, rng_w;
    int32_t d2ok;         // 8-bit filter may refine "comparison 0 passes" candidates with comparison 1 (see filter_lane)
    // Differences (mod 2^(8W)) of comparison 0 that do NOT lead to a J0 advance, stored as the
    // SWAR constant the filter kernel consumes:  W=1: k * 0x01010101 ;  W=2: ((1-k) & 0xFFFF) * 0x00010001
    uint32_t keys[MMG_MAXL + 1];
    // 16-bit pre-filter constants: upper half (1 - k), lower half (-k)  (scan_kernels.cu prefilter16);
    // 8-bit, odd lags: the key constants of the difference registers whose PREVIOUS pair carries the bias (filter8)
    uint32_t pkeys[MMG_MAXL + 1];
    MmgCheck chk[MMG_MAXL];
    int32_t tab_key[MMG_MAXL];     // exact signed difference
    int32_t tab_val[MMG_MAXL];     // skip, already max(.,1)
};

// Keywords longer than MMG_MAXL (the reference accepts any length): the same program with its per-position arrays in
// device memory instead of the kernel-parameter block.  Only the per-chain kernels (g_walk_long) run it -- one thread
// per (block, alignment) chain, exactly the CPU's loop; correct for every length, nowhere near the streaming path's speed.
#define MMG_MAXL_LONG 32767         // (keyword indices are int16_t in MmgCheck)
struct MmgLongProgram {
    int32_t W, L, modular, ncheck, ntab, tab_default, match_jump, first_lit, opp_idx;
    const MmgCheck *chk;            // [ncheck] device
    const int32_t *tab_key;         // [ntab] device
    const int32_t *tab_val;
};

#endif
