// search_engine.hpp -- mmoore::SearchEngine<T>: file -> blocks -> GPU scan -> sorted results.
//
// Source-compatible with the reference's include/mmoore/search_engine.hpp:16-58: SearchResult is
// an aggregate {offset, values_map, preview}; SearchConfig has the same fields and defaults;
// SearchStep is an unscoped enum; run() keeps the callback protocol ((0,Initializing),
// (0,Searching), one (pct,Searching) per block, (100,GeneratingPreviews)), the abort contract
// (return {} once the flag is raised) and the std::runtime_error("File not found").
// preferred_num_threads is advisory here (callbacks come from the calling thread only);
// preferred_search_block_size stays binding -- it defines where the skip chains restart.
#ifndef MMOORE_B200_SEARCH_ENGINE_HPP
#define MMOORE_B200_SEARCH_ENGINE_HPP

#include <atomic>
#include <filesystem>
#include <fstream>
#include <functional>
#include <thread>
#include <vector>

#include "mmoore/byteswap.hpp"
#include "mmoore/monkey_moore.hpp"

namespace mmoore {

template <typename DataType>
struct SearchResult {
   uint64_t offset;
   typename MonkeyMoore<DataType>::equivalency_map values_map;
   std::string preview;
};

struct SearchConfig {
   std::filesystem::path file_path;

   bool is_relative_search = true;
   mmoore::Endianness endianness = Endianness::Little;

   std::vector<CharType> keyword;
   std::vector<CharType> custom_char_seq = {};
   CharType wildcard = '*';

   std::vector<short> reference_values = {};

   int preferred_num_threads = std::thread::hardware_concurrency();
   int preferred_search_block_size = 524288;
   int preferred_preview_width = 50;
};

enum SearchStep { Initializing, Searching, GeneratingPreviews, Aborting };

template <typename DataType>
class SearchEngine {
public:
   using ProgressCallback = std::function<void(int, const SearchStep)>;

   explicit SearchEngine(const SearchConfig &cfg) : config(cfg) {}

   std::vector<SearchResult<DataType>> run(ProgressCallback on_progress, std::atomic<bool> &abort_flag,
                                           bool generate_previews = false);

private:
   SearchConfig config;

   std::string generate_preview(std::ifstream &file, uint64_t file_size, uint64_t match_offset,
                                std::map<CharType, DataType> &values_map);
   std::string decode_raw_data(std::map<CharType, DataType> &values_map, std::vector<DataType> &raw_data);
};

}  // namespace mmoore

#endif
