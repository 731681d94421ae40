// pattern.cpp -- host-side pattern compiler.
//
// Turns what the reference's MonkeyMoore<Ty> constructors compute
// (/root/reference/src/core/monkey_moore.cpp:12-39 constructors, :54-78 initialize,
// :83-100 preprocess, :106-142 preprocess_no_wildcards, :144-304 preprocess_with_wildcards)
// into the MmgProgram POD the kernels consume.  Written from the behavioural spec in
// SURVEY.md Appendix A; it shares no code with oracle/ (the checker) by design.
#include "pattern.hpp"

#include "../../include/mmoore_b200.h"

#include <algorithm>
#include <cmath>
#include <climits>
#include <cstring>
#include <memory>

namespace {

inline bool is_upper(uint32_t c) { return c >= 'A' && c <= 'Z'; }   // include/mmoore/text_utils.hpp:41-43
inline bool is_lower(uint32_t c) { return c >= 'a' && c <= 'z'; }   // include/mmoore/text_utils.hpp:48-50

// The reference subtracts char32_t code points (or int indices) and stores the result in an
// int, i.e. arithmetic modulo 2^32 (src/core/monkey_moore.cpp:236, 560-563, 578-581).
inline int32_t wrap_diff(int64_t a, int64_t b) { return static_cast<int32_t>(static_cast<uint32_t>(a - b)); }

struct SparseSkip {
    std::vector<int32_t> key, val;
    int32_t fallback = 0;
    int find(int32_t k) const {
        for (size_t j = 0; j < key.size(); j++)
            if (key[j] == k) return static_cast<int>(j);
        return -1;
    }
};

}  // namespace

int mmg_program::value_of(uint32_t c) const {
    if (char_seq.empty()) return static_cast<int>(c);
    auto it = seq_index.find(c);
    return it == seq_index.end() ? 0 : it->second;   // operator[] default-inserts 0
}

int mmg_compile_pattern(const uint32_t *keyword, int keyword_len, uint32_t wildcard, const uint32_t *char_seq,
                        int char_seq_len, bool value_scan, int elem_bits, mmg_program **out, std::string &err) {
    *out = nullptr;
    if (elem_bits != 8 && elem_bits != 16) { err = "elem_bits must be 8 or 16"; return MMG_ERR_ARG; }
    if (keyword_len <= 0 || keyword == nullptr) { err = "empty keyword"; return MMG_ERR_EMPTY; }
    if (keyword_len > MMG_MAXL_LONG) { err = "keyword longer than 32767 elements"; return MMG_ERR_TOO_LONG; }
    if (char_seq_len < 0 || (char_seq_len > 0 && char_seq == nullptr)) { err = "bad char_seq"; return MMG_ERR_ARG; }

    auto prog = new mmg_program();
    std::unique_ptr<mmg_program> guard(prog);
    const int L = keyword_len;
    const int64_t vmax = elem_bits == 8 ? 255 : 65535;
    const uint32_t vmask = static_cast<uint32_t>(vmax);

    prog->elem_bits = elem_bits;
    prog->keyword.assign(keyword, keyword + L);
    prog->normalized = prog->keyword;
    prog->wildcard = wildcard;
    if (char_seq_len > 0) prog->char_seq.assign(char_seq, char_seq + char_seq_len);
    for (int i = 0; i < char_seq_len; i++) prog->seq_index[char_seq[i]] = i;   // later duplicates win (:87-88)

    // ---- mode selection (initialize, :66-77)
    bool has_wildcards = std::count(prog->keyword.begin(), prog->keyword.end(), wildcard) > 0;
    int uppers = 0, lowers = 0;
    for (uint32_t c : prog->keyword) { uppers += is_upper(c); lowers += is_lower(c); }
    prog->has_case_change = prog->char_seq.empty() && !value_scan && uppers > 0 && lowers > 0;
    prog->mode = value_scan ? 2 : ((has_wildcards || prog->has_case_change) ? 1 : 0);

    std::vector<int32_t> ed(L, 0);        // expected difference at keyword index i
    std::vector<int> prev(L, 0);          // index compared against
    std::vector<bool> literal(L, true);
    std::vector<int32_t> cap(L, INT32_MAX);
    SparseSkip skip;
    auto in_table = [&](int32_t d) {      // 0 <= d + max < 2*(max+1)   (:128-130, :259-261)
        int64_t index = static_cast<int64_t>(d) + vmax;
        return index >= 0 && index < 2 * (vmax + 1);
    };

    int lead = 0, first_lit = 0, opp_idx = -1;

    if (prog->mode != 1) {
        // ---- simple relative / value scan
        for (int i = 0; i < L; i++) {
            prev[i] = i == 0 ? L - 1 : i - 1;
            ed[i] = wrap_diff(prog->value_of(prog->keyword[i]), prog->value_of(prog->keyword[prev[i]]));
        }
        skip.fallback = L - 1;
        for (int i = L - 1; i >= 0; i--) {
            if (!in_table(ed[i])) { err = "Skip table index out of bounds"; return MMG_ERR_SKIP_OOB; }
            // a slot is written only while it still holds the default: the rightmost occurrence wins (:133-135)
            if (skip.find(ed[i]) < 0) { skip.key.push_back(ed[i]); skip.val.push_back(L - 1 - i); }
        }
    } else {
        // ---- wildcard relative (explicit wildcards and/or mixed-case keywords)
        if (prog->char_seq.empty()) {
            prog->mostly_lowercase = lowers > uppers;                           // :163
            if (uppers > 0 && lowers > 0) {                                     // :165-180
                const bool drop_lower = uppers > lowers;                        // ties drop the UPPERCASE letters
                for (auto &c : prog->normalized)
                    if (drop_lower ? is_lower(c) : is_upper(c)) c = wildcard;
            }
        }
        const auto &nk = prog->normalized;
        int last_lit = -1;
        first_lit = L;
        for (int i = 0; i < L; i++) {
            literal[i] = nk[i] != wildcard;
            if (literal[i]) { if (first_lit == L) first_lit = i; last_lit = i; }
            prev[i] = i;
        }
        // bridge literal -> previous literal; the first literal wraps to the last (:222-247)
        int bridge = last_lit;
        for (int i = 0; i < L; i++) {
            if (!literal[i]) continue;
            prev[i] = bridge;
            ed[i] = wrap_diff(prog->value_of(nk[i]), prog->value_of(nk[bridge]));
            bridge = i;
        }
        // skip table written through `char`; i descends to 1 and later writes override (:249-276)
        skip.fallback = static_cast<int32_t>(static_cast<signed char>(L - 1));
        for (int i = L - 1; i > 0; --i) {
            if (!in_table(ed[i])) { err = "Skip table index out of bounds"; return MMG_ERR_SKIP_OOB; }
            int wc_after = static_cast<int>(std::count(nk.begin() + i + 1, nk.end(), wildcard));
            int32_t v = static_cast<int32_t>(static_cast<signed char>(L - wc_after - i - 1));
            int j = skip.find(ed[i]);
            if (j < 0) { skip.key.push_back(ed[i]); skip.val.push_back(v); }
            else skip.val[j] = v;
        }
        // wildcard skip table (:278-303)
        for (int i = 0; i < L; i++) {
            if (!literal[i]) { cap[i] = 1; continue; }
            int last_wc = 0;   // "not found" is treated as index 0
            for (int j = 0; j < i; j++) if (nk[j] == wildcard) last_wc = j;
            cap[i] = static_cast<unsigned char>(std::max(i - last_wc - 1, 1));
        }
        while (lead < L && nk[lead] == wildcard) lead++;                       // :438-441
        if (prog->has_case_change) {                                            // :490-499
            for (int i = 0; i < L; i++) {
                bool target = prog->mostly_lowercase ? is_upper(prog->keyword[i]) : is_lower(prog->keyword[i]);
                if (target) { opp_idx = i; break; }
            }
        }
    }

    // A non-positive advance after a match never terminates in the reference (:398, :526).
    const int match_jump = L - 1 - lead;
    if (match_jump < 1) { err = "pattern never advances after a match (keyword too short / all wildcards)"; return MMG_ERR_HANG; }

    // ---- emit the device program
    MmgProgram &d = prog->dev;
    std::memset(&d, 0, sizeof(d));
    d.W = elem_bits / 8;
    d.L = L;
    d.modular = prog->mode == 1;
    d.match_jump = match_jump;
    d.first_lit = first_lit < L ? first_lit : 0;
    d.opp_idx = opp_idx;
    d.tab_default = std::max(skip.fallback, 1);
    if (L > MMG_MAXL) {
        // long keyword: the arrays go to device memory (MmgLongProgram), evaluated by the per-chain kernels only
        prog->is_long = true;
        for (size_t j = 0; j < skip.key.size(); j++) {
            prog->long_tab_key.push_back(skip.key[j]);
            prog->long_tab_val.push_back(std::max(skip.val[j], 1));
        }
        int first_literal = -1;
        for (int i = 0; i < L; i++) if (literal[i]) { first_literal = i; break; }
        int32_t jmax = match_jump;
        for (int i = L - 1; i >= 0; i--) {
            if (!literal[i] || i == first_literal) continue;
            MmgCheck c;
            c.i = static_cast<int16_t>(i);
            c.lag = static_cast<int16_t>(i - prev[i]);
            c.ed = ed[i];
            c.cap = cap[i];
            prog->long_chk.push_back(c);
        }
        d.ntab = 0;
        d.ncheck = 0;
        d.nkeys = -1;
        d.J0 = 1;
        d.Jmax = std::max(jmax, 1);
        *out = guard.release();
        return MMG_OK;
    }
    d.ntab = static_cast<int32_t>(skip.key.size());
    int32_t max_skip = d.tab_default;
    for (int j = 0; j < d.ntab; j++) {
        d.tab_key[j] = skip.key[j];
        d.tab_val[j] = std::max(skip.val[j], 1);
        max_skip = std::max(max_skip, d.tab_val[j]);
    }
    // comparisons in evaluation order: literals right to left, without the first literal
    bool seen_first = false;
    int first_literal_index = -1;
    for (int i = 0; i < L; i++) if (literal[i]) { first_literal_index = i; break; }
    (void)seen_first;
    d.ncheck = 0;
    int32_t jmax = match_jump;
    for (int i = L - 1; i >= 0; i--) {
        if (!literal[i] || i == first_literal_index) continue;
        MmgCheck &c = d.chk[d.ncheck++];
        c.i = static_cast<int16_t>(i);
        c.lag = static_cast<int16_t>(i - prev[i]);
        c.ed = ed[i];
        c.cap = cap[i];
        jmax = std::max(jmax, std::min(c.cap, max_skip));
    }
    d.Jmax = jmax;

    if (d.ncheck == 0) {
        // a single literal: every window matches
        d.nkeys = -1;
        d.J0 = match_jump;
    } else {
        const int32_t cap0 = d.chk[0].cap;
        d.J0 = std::min(cap0, d.tab_default);
        if (d.W == 1) {
            // 8-bit filter keys are EXACT signed differences (scan_kernels.cu filter8).  keys[0] is the difference
            // comparison 0 expects; in wildcard mode the comparison is modulo 256, so both signed values of that
            // residue pass.  The other keys are the table entries whose advance differs from J0.
            std::vector<int32_t> keys;
            auto add_key = [&](int32_t diff) {
                if (diff > vmax || diff < -vmax) return;                       // a difference the data cannot produce
                if (std::find(keys.begin(), keys.end(), diff) == keys.end()) keys.push_back(diff);
            };
            const int32_t ed0 = d.chk[0].ed;
            if (d.modular) {
                const int32_t r = static_cast<int32_t>(static_cast<uint32_t>(ed0) & vmask);
                add_key(r);
                if (r != 0) add_key(r - 256);
            } else {
                add_key(ed0);
            }
            const bool pass_first = !keys.empty() && keys[0] == ed0;
            for (int j = 0; j < d.ntab; j++)
                if (std::min(cap0, d.tab_val[j]) != d.J0) add_key(d.tab_key[j]);
            d.nkeys = static_cast<int32_t>(keys.size());
            if (d.nkeys == 0) { keys.push_back(1 << 20); d.nkeys = 1; }        // nothing can ever be flagged: one impossible key
            // depth-2 refinement: comparisons 0 and 1 look at adjacent element pairs with exact arithmetic
            d.d2ok = (!d.modular && d.ncheck >= 2 && d.chk[0].lag == 1 && d.chk[1].lag == 1 &&
                      d.chk[1].i == d.chk[0].i - 1 && pass_first) ? 1 : 0;
            for (size_t j = 0; j < keys.size(); j++) {
                const uint32_t k = static_cast<uint32_t>(keys[j]);
                // difference registers with a biased current pair hold 256 + d in both halves ...
                d.keys[j] = ((0x10000u - 256u - k) & 0xFFFFu) * 0x00010001u;
                // ... those with a biased PREVIOUS pair (odd lags, scan_kernels.cu filter8) hold d - 256 in the low half
                // and d - 257 in the high half (the low half always borrows)
                d.pkeys[j] = (((257u - k) & 0xFFFFu) << 16) | ((256u - k) & 0xFFFFu);
            }
        } else {
            std::vector<uint32_t> keys;
            auto add_key = [&](int32_t diff) {
                if (!d.modular && (diff > vmax || diff < -vmax)) return;   // an exact difference the data cannot produce
                uint32_t k = static_cast<uint32_t>(diff) & vmask;
                if (std::find(keys.begin(), keys.end(), k) == keys.end()) keys.push_back(k);
            };
            add_key(d.chk[0].ed);
            for (int j = 0; j < d.ntab; j++)
                if (std::min(cap0, d.tab_val[j]) != d.J0) add_key(d.tab_key[j]);
            d.nkeys = static_cast<int32_t>(keys.size());
            for (size_t j = 0; j < keys.size(); j++) {
                d.keys[j] = ((1u - keys[j]) & 0xFFFFu) * 0x00010001u;
                d.pkeys[j] = (((1u - keys[j]) & 0xFFFFu) << 16) | ((0u - keys[j]) & 0xFFFFu);
            }
            // range stage of the pre-filter: the shortest arc of the 16-bit circle that holds every key.  Worth its
            // 8 instructions per row when a 512-position row rarely holds a difference inside the arc.
            d.rng_w = 0xFFFFFFFFu;
            d.rng_c = 0;
            if (keys.size() >= 2) {
                std::vector<uint32_t> sorted(keys);
                std::sort(sorted.begin(), sorted.end());
                uint32_t best_gap = sorted[0] + 0x10000u - sorted.back(), lo = sorted[0];
                for (size_t j = 1; j < sorted.size(); j++)
                    if (sorted[j] - sorted[j - 1] > best_gap) { best_gap = sorted[j] - sorted[j - 1]; lo = sorted[j]; }
                const uint32_t w = 0x10000u - best_gap;                       // (k - lo) mod 2^16 <= w for every key
                const double p_row = 1.0 - std::pow(1.0 - (w + 2.0) / 65536.0, 512.0);
                if (p_row < 0.9 - 1.0 / static_cast<double>(keys.size())) {
                    d.rng_w = w;
                    d.rng_c = (((1u - lo) & 0xFFFFu) << 16) | ((0u - lo) & 0xFFFFu);
                }
            }
        }
    }

    *out = guard.release();
    return MMG_OK;
}
