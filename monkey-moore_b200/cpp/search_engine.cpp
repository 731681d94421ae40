// search_engine.cpp -- mmoore::SearchEngine<T> over the C-ABI (include/mmoore_b200.h).
//
// Replaces the reference's src/core/search_engine.cpp.  What is gone: one ifstream + one
// std::async thread per block, a full buffer copy per alignment, the in-place byte swap and the
// dispatcher's 5 ms polling sleep (:104-188).  What replaces it: the file is read in large slabs
// of whole blocks into page-locked memory and each slab is handed to mmg_engine_scan, which runs
// every (block, alignment) chain on the GPU in one pass and returns offsets already sorted.
// What is kept exactly: block geometry (:218-253), the callback protocol and abort contract
// (:47, :80, :161-165, :177-187, :191), "File not found" (:43-45), result order (:193-197) and the
// preview text (:256-348, including the sticky stream state of the shared ifstream).
#include "mmoore/search_engine.hpp"

#include "mmoore_b200.h"

#include <cmath>
#include <cstring>
#include <iomanip>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace {

[[noreturn]] void throw_last(int rc) {
   if (rc == MMG_ERR_SKIP_OOB) throw std::runtime_error("Skip table index out of bounds");
   const char *msg = mmg_last_error();
   throw std::runtime_error(msg && *msg ? msg : "monkey-moore GPU search failed");
}

// UTF-8 encoding of one code point (the reference goes through std::codecvt_utf8<char32_t>,
// src/core/encoding.hpp:25-28)
std::string to_utf8(char32_t cp) {
   std::string out;
   if (cp < 0x80) {
      out += static_cast<char>(cp);
   } else if (cp < 0x800) {
      out += static_cast<char>(0xC0 | (cp >> 6));
      out += static_cast<char>(0x80 | (cp & 0x3F));
   } else if (cp < 0x10000) {
      out += static_cast<char>(0xE0 | (cp >> 12));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
   } else {
      out += static_cast<char>(0xF0 | (cp >> 18));
      out += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
   }
   return out;
}

struct PinnedBuffer {
   uint8_t *ptr = nullptr;
   bool pinned = false;
   explicit PinnedBuffer(uint64_t n) {
      ptr = static_cast<uint8_t *>(mmg_host_alloc(n));
      pinned = ptr != nullptr;
      if (!ptr) ptr = static_cast<uint8_t *>(::operator new(n ? n : 1));
   }
   ~PinnedBuffer() {
      if (pinned) mmg_host_free(ptr);
      else ::operator delete(ptr);
   }
};

constexpr uint64_t kSlabBytes = 256ull << 20;   // bytes of whole blocks handed to one GPU scan

}  // namespace

template <typename DataType>
std::vector<mmoore::SearchResult<DataType>> mmoore::SearchEngine<DataType>::run(ProgressCallback on_progress,
                                                                                 std::atomic<bool> &abort_flag,
                                                                                 bool generate_previews) {
   std::vector<mmoore::SearchResult<DataType>> results;

   if (!std::filesystem::exists(config.file_path)) throw std::runtime_error("File not found");
   on_progress(0, mmoore::SearchStep::Initializing);
   const uint64_t file_size = std::filesystem::file_size(config.file_path);

   std::unique_ptr<MonkeyMoore<DataType>> searcher;
   if (config.is_relative_search)
      searcher = std::make_unique<MonkeyMoore<DataType>>(config.keyword, config.wildcard, config.custom_char_seq);
   else
      searcher = std::make_unique<MonkeyMoore<DataType>>(config.reference_values);

   const size_t pattern_len = config.is_relative_search ? config.keyword.size() : config.reference_values.size();
   const uint64_t overlap = (pattern_len - 1) * sizeof(DataType);
   const uint32_t block = static_cast<uint32_t>(config.preferred_search_block_size);
   const uint64_t num_blocks = mmg_num_blocks(file_size, block);

   float total_progress = 0.0f;
   const float progress_increment = 100.0f / num_blocks;

   on_progress(0, mmoore::SearchStep::Searching);

   if (num_blocks > 0) {
      std::ifstream file(config.file_path, std::ios::binary);
      if (!file.is_open()) throw std::runtime_error("Worker thread failed to open file: " + config.file_path.string());

      const uint64_t blocks_per_slab = std::max<uint64_t>(1, kSlabBytes / block);
      const uint64_t slab_capacity = std::min<uint64_t>(file_size, blocks_per_slab * block + overlap);
      PinnedBuffer slab(slab_capacity);
      const bool big_endian = config.endianness == mmoore::Endianness::Big;

      for (uint64_t first = 0; first < num_blocks; first += blocks_per_slab) {
         const uint64_t n = std::min<uint64_t>(blocks_per_slab, num_blocks - first);
         const uint64_t lo = first * block;
         const uint64_t hi = std::min<uint64_t>(file_size, (first + n) * static_cast<uint64_t>(block) + overlap);
         file.clear();
         file.seekg(static_cast<std::streamoff>(lo));
         file.read(reinterpret_cast<char *>(slab.ptr), static_cast<std::streamsize>(hi - lo));
         const uint64_t got = static_cast<uint64_t>(file.gcount());
         if (got < hi - lo) std::memset(slab.ptr + got, 0, hi - lo - got);   // file shrank underneath us

         mmg_results *res = nullptr;
         const int rc = mmg_engine_scan(searcher->program(), slab.ptr, hi - lo, MMG_MEM_HOST, file_size, block, first, n,
                                        big_endian ? 1 : 0, &res);
         if (rc != MMG_OK) throw_last(rc);
         const uint64_t count = mmg_results_count(res);
         if (count) {
            std::vector<uint64_t> offsets(count);
            std::vector<uint32_t> values(2 * count);
            const int rc2 = mmg_results_copy(res, 0, count, offsets.data(), values.data());
            if (rc2 != MMG_OK) { mmg_results_free(res); throw_last(rc2); }
            results.reserve(results.size() + count);
            for (uint64_t i = 0; i < count; i++)
               results.push_back({offsets[i], searcher->table_from_values(values[2 * i], values[2 * i + 1]), std::string()});
         }
         mmg_results_free(res);

         // one callback per block, exactly as the reference's workers report (float accumulation included)
         for (uint64_t b = 0; b < n; b++) {
            total_progress += progress_increment;
            on_progress(static_cast<int>(total_progress), SearchStep::Searching);
            if (abort_flag) return {};
         }
      }
   }
   if (abort_flag) return {};

   on_progress(100, GeneratingPreviews);
   // slabs arrive in file order and each scan returns ascending offsets: already sorted

   if (generate_previews && !results.empty()) {
      std::ifstream preview_file(config.file_path, std::ios::binary);
      if (!preview_file.is_open())
         throw std::runtime_error("Failed to open file to generate previews: " + config.file_path.string());
      for (auto &result : results) result.preview = generate_preview(preview_file, file_size, result.offset, result.values_map);
   }
   return results;
}

template <typename DataType>
std::string mmoore::SearchEngine<DataType>::generate_preview(std::ifstream &file, uint64_t file_size, uint64_t match_offset,
                                                             std::map<CharType, DataType> &values_map) {
   const int64_t elem = static_cast<int64_t>(sizeof(DataType));
   const int64_t width = config.preferred_preview_width;
   // centre the match in the window
   const int64_t keyword_half = static_cast<int64_t>(config.keyword.size() / 2);
   const int64_t window_half = width / 2;
   int64_t back = (window_half - keyword_half) * elem;
   back = (back + (elem - 1)) & ~(elem - 1);                       // keep multi-byte elements aligned
   int64_t start = static_cast<int64_t>(match_offset) - back;
   const int64_t end = start + width * elem;
   if (static_cast<uint64_t>(end) > file_size) start -= static_cast<int64_t>(static_cast<uint64_t>(end) - file_size);

   // the stream is shared by all previews and deliberately NOT cleared: a short read leaves it failed
   file.seekg(std::max<int64_t>(0, start), std::ios::beg);
   std::vector<DataType> window(static_cast<size_t>(width));
   file.read(reinterpret_cast<char *>(window.data()), width * elem);
   window.resize(static_cast<size_t>(file.gcount()) / sizeof(DataType));
   if (sizeof(DataType) > 1) mmoore::adjust_endianness(window.data(), window.size(), config.endianness);
   return decode_raw_data(values_map, window);
}

template <typename DataType>
std::string mmoore::SearchEngine<DataType>::decode_raw_data(std::map<CharType, DataType> &values_map,
                                                            std::vector<DataType> &raw_data) {
   std::ostringstream text;
   if (!config.is_relative_search) {
      // value scan: hex dump of the window
      text << std::hex << std::uppercase << std::setfill('0');
      for (size_t i = 0; i < raw_data.size(); ++i) {
         if (i) text << " ";
         text << std::setw(sizeof(DataType) * 2) << static_cast<uint64_t>(raw_data[i]);
      }
      return text.str();
   }
   const bool ascii_search = config.custom_char_seq.empty();
   std::unordered_map<DataType, std::string> glyph;
   glyph.reserve(values_map.size() * (ascii_search ? 26 : 1));
   for (const auto &[character, value] : values_map) {             // ascending code point; later entries win
      if (ascii_search && (character == U'a' || character == U'A')) {
         for (int k = 0; k < 26; ++k)
            glyph[static_cast<DataType>(value + static_cast<DataType>(k))] = to_utf8(character + static_cast<char32_t>(k));
      } else {
         glyph[value] = to_utf8(character);
      }
   }
   for (const DataType v : raw_data) {
      const auto it = glyph.find(v);
      if (it != glyph.end()) text << it->second;
      else text << "#";
   }
   return text.str();
}

template class mmoore::SearchEngine<uint8_t>;
template class mmoore::SearchEngine<uint16_t>;
