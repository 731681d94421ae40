// pattern.hpp -- host-side pattern compiler: keyword / value list -> MmgProgram.
#ifndef MMG_PATTERN_HPP
#define MMG_PATTERN_HPP

#include "program.h"

#include <atomic>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

// What mmg_program (the C-ABI handle) points at.
struct mmg_program {
    MmgProgram dev;                        // what the kernels consume

    int elem_bits = 8;
    int mode = 0;                          // 0 simple_relative, 1 wildcard_relative, 2 value_scan
    std::vector<uint32_t> keyword;         // as given
    std::vector<uint32_t> normalized;      // minority-case letters replaced by the wildcard (wildcard mode)
    uint32_t wildcard = 0;
    std::vector<uint32_t> char_seq;
    std::map<uint32_t, int> seq_index;     // char -> position; unknown chars read as 0 (reference: operator[])
    bool has_case_change = false;
    bool mostly_lowercase = false;

    // sizing hints remembered from the previous scan with this pattern (not part of its semantics)
    mutable std::atomic<uint64_t> last_count{0};
    mutable std::atomic<uint64_t> last_events_per_warp{0};
    mutable std::atomic<uint64_t> last_events{0};          // total events of the previous scan
    mutable std::atomic<uint64_t> last_bytes{0};           // bytes the previous scan covered (0: no scan yet)

    // keyword longer than MMG_MAXL: dev holds the scalars only, the arrays live here (and, from the first scan on, in
    // device memory: capi.cu uploads them per device)
    bool is_long = false;
    std::vector<MmgCheck> long_chk;
    std::vector<int32_t> long_tab_key, long_tab_val;

    int value_of(uint32_t c) const;        // code point, or index in char_seq (0 when absent)
};

// Returns MMG_OK or an MMG_ERR_* code; on success *out is a heap object (delete to free).
int mmg_compile_pattern(const uint32_t *keyword, int keyword_len, uint32_t wildcard, const uint32_t *char_seq,
                        int char_seq_len, bool value_scan, int elem_bits, mmg_program **out, std::string &err);

#endif
