// monkey_moore.cpp -- MonkeyMoore<Ty> over the C-ABI (include/mmoore_b200.h).
//
// Replaces the reference's src/core/monkey_moore.cpp: the constructors (:12-39) become
// mmg_program_create_*, search() (:41-49, 316-546) becomes mmg_search on the GPU, and the
// per-match equivalency_map construction (:374-393, :472-521) becomes mmg_program_table applied to
// the two raw element values the device reports for each match.
#include "mmoore/monkey_moore.hpp"

#include "mmoore_b200.h"

#include <stdexcept>

namespace {

[[noreturn]] void throw_for(int rc) {
   // the reference's own message for the only exception its constructors raise
   if (rc == MMG_ERR_SKIP_OOB) throw std::runtime_error("Skip table index out of bounds");
   const char *msg = mmg_last_error();
   throw std::runtime_error(msg && *msg ? msg : "monkey-moore GPU search failed");
}

template <class Ty> constexpr int elem_bits() { return static_cast<int>(sizeof(Ty) * 8); }

}  // namespace

template <class Ty>
MonkeyMoore<Ty>::MonkeyMoore(const std::vector<CharType> &keyword, CharType wildcard, const std::vector<CharType> &char_seq) {
   assert(!keyword.empty());
   static_assert(sizeof(CharType) == sizeof(uint32_t), "CharType is UTF-32");
   const int rc = mmg_program_create_keyword(reinterpret_cast<const uint32_t *>(keyword.data()), static_cast<int>(keyword.size()),
                                             static_cast<uint32_t>(wildcard),
                                             reinterpret_cast<const uint32_t *>(char_seq.data()), static_cast<int>(char_seq.size()),
                                             elem_bits<Ty>(), &handle);
   if (rc != MMG_OK) throw_for(rc);
}

template <class Ty>
MonkeyMoore<Ty>::MonkeyMoore(const std::vector<short> &reference_values) {
   assert(!reference_values.empty());
   static_assert(sizeof(short) == sizeof(int16_t), "short is 16 bit");
   const int rc = mmg_program_create_values(reinterpret_cast<const int16_t *>(reference_values.data()),
                                            static_cast<int>(reference_values.size()), elem_bits<Ty>(), &handle);
   if (rc != MMG_OK) throw_for(rc);
}

template <class Ty>
MonkeyMoore<Ty>::~MonkeyMoore() {
   mmg_program_free(handle);
}

template <class Ty>
typename MonkeyMoore<Ty>::equivalency_map MonkeyMoore<Ty>::table_from_values(uint32_t v0, uint32_t v1) const {
   equivalency_map table;
   const int n = mmg_program_table_size(handle);
   if (n == 0) return table;
   std::vector<uint32_t> keys(n), values(n);
   mmg_program_table(handle, v0, v1, keys.data(), values.data());
   auto hint = table.end();
   for (int i = 0; i < n; i++)   // keys arrive in ascending order
      hint = table.emplace_hint(hint, static_cast<CharType>(keys[i]), static_cast<Ty>(values[i]));
   return table;
}

template <class Ty>
std::vector<typename MonkeyMoore<Ty>::result_type> MonkeyMoore<Ty>::search(const Ty *data, uint64_t data_len) {
   std::vector<result_type> results;
   mmg_results *res = nullptr;
   const int rc = mmg_search(handle, data, data_len, MMG_MEM_HOST, &res);
   if (rc != MMG_OK) throw_for(rc);
   const uint64_t n = mmg_results_count(res);
   if (n) {
      std::vector<uint64_t> positions(n);
      std::vector<uint32_t> values(2 * n);
      const int rc2 = mmg_results_copy(res, 0, n, positions.data(), values.data());
      if (rc2 != MMG_OK) { mmg_results_free(res); throw_for(rc2); }
      results.reserve(n);
      // consecutive matches very often share their base values: build each distinct table once
      uint32_t last0 = 0, last1 = 0;
      bool have = false;
      equivalency_map table;
      for (uint64_t i = 0; i < n; i++) {
         if (!have || values[2 * i] != last0 || values[2 * i + 1] != last1) {
            table = table_from_values(values[2 * i], values[2 * i + 1]);
            last0 = values[2 * i]; last1 = values[2 * i + 1]; have = true;
         }
         results.emplace_back(positions[i], table);
      }
   }
   mmg_results_free(res);
   return results;
}

template class MonkeyMoore<uint8_t>;
template class MonkeyMoore<uint16_t>;
