// monkey_moore.hpp -- MonkeyMoore<Ty>: pattern setup + relative search, executed on the GPU.
//
// Source-compatible with the reference's include/mmoore/monkey_moore.hpp:16-51 (global CharType,
// MonkeyMoore<Ty> with equivalency_map / result_type, the keyword and value-scan constructors and
// search()).  The object owns a compiled pattern handle of the C-ABI (include/mmoore_b200.h);
// search() copies the caller's elements to the device, replays the reference's skip chain there
// and returns the same (position, inferred table) pairs in the same order.  Errors surface as the
// reference's exceptions: std::runtime_error("Skip table index out of bounds") from the
// constructors.  Deviation (documented in DESIGN.md): patterns whose match advance is <= 0 hang
// the reference; here the constructor throws std::runtime_error instead.
#ifndef MMOORE_B200_MONKEY_MOORE_HPP
#define MMOORE_B200_MONKEY_MOORE_HPP

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

using CharType = char32_t;

struct mmg_program;   // include/mmoore_b200.h

template <class Ty>
class MonkeyMoore {
public:
   using equivalency_map = std::map<CharType, Ty>;
   using result_type = std::pair<uint64_t, equivalency_map>;

   // Relative search for `keyword`; `wildcard` marks don't-care positions; `char_seq` is an
   // optional user-defined character order (e.g. a Kana table).
   MonkeyMoore(const std::vector<CharType> &keyword, CharType wildcard = 0, const std::vector<CharType> &char_seq = {});

   // Value scan: the relative pattern of a raw numeric sequence.
   MonkeyMoore(const std::vector<short> &reference_values);

   ~MonkeyMoore();
   MonkeyMoore(const MonkeyMoore &) = delete;
   MonkeyMoore &operator=(const MonkeyMoore &) = delete;

   // data_len is in elements; results are ascending by position.
   std::vector<result_type> search(const Ty *data, uint64_t data_len);

   // GPU-side access for mmoore::SearchEngine (not part of the reference surface).
   const mmg_program *program() const { return handle; }
   equivalency_map table_from_values(uint32_t v0, uint32_t v1) const;

private:
   mmg_program *handle = nullptr;
};

#endif
