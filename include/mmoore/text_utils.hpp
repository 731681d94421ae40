// text_utils.hpp -- small sequence / ASCII helpers of the search API.
//
// Same public surface as the reference's include/mmoore/text_utils.hpp:14-56; pinned by the
// reference's tests/test_text_utils.cpp.
#ifndef MMOORE_B200_TEXT_UTILS_HPP
#define MMOORE_B200_TEXT_UTILS_HPP

#include <cstdint>
#include <iterator>

// Index of the last element equal to `v` in [start, end), or -1.
// The pattern compiler (monkey-moore_b200/csrc/pattern.cpp) restates the reference's use of it: the distance from a
// keyword position back to the previous wildcard bounds the advance after a mismatch at that position.
template <class FwdIt, class T>
inline int find_last_index(FwdIt start, const FwdIt end, const T &v) {
   int found = -1;
   int index = 0;
   while (start != end) {
      if (*start == v) found = index;
      ++start;
      ++index;
   }
   return found;
}

// Length of the run of elements equal to `v` at the beginning of [start, end).
// Leading wildcards shorten the advance after a match (keyword length - 1 - leading wildcards); a keyword that would
// not advance at all is rejected by this library instead of looping forever.
template <class FwdIt, class T>
inline int count_prefix_length(FwdIt start, const FwdIt end, const T &v) {
   int run = 0;
   while (start != end && *start == v) {
      ++run;
      ++start;
   }
   return run;
}

// Code-point range checks (no locale, no <cctype>): char32_t values above 127 are never ASCII letters or digits.
inline bool is_ascii_upper(const char32_t &c) { return c >= U'A' && c <= U'Z'; }
inline bool is_ascii_lower(const char32_t &c) { return c >= U'a' && c <= U'z'; }
inline bool is_ascii_digit(const char32_t &c) { return c >= U'0' && c <= U'9'; }

#endif
