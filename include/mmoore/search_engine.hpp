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

// One match.  Aggregate with exactly this member order: the reference's tests brace-initialise it as
// {offset, {}, ""} (tests/test_search_engine.cpp:47).
template <typename DataType>
struct SearchResult {
   uint64_t offset;                                                // byte offset of the match in the file
   typename MonkeyMoore<DataType>::equivalency_map values_map;     // inferred character -> value table (empty for a value scan)
   std::string preview;                                            // decoded window around the match, when requested
};

// What to search for and how.  Field names, order and defaults are the reference's
// (include/mmoore/search_engine.hpp:23-38); the GUI fills this struct (src/gui/monkey_frame.cpp:555-571).
struct SearchConfig {
   std::filesystem::path file_path;                                // throws std::runtime_error("File not found") from run()

   bool is_relative_search = true;                                 // false: value scan over reference_values
   mmoore::Endianness endianness = Endianness::Little;             // byte order of 16-bit elements in the file

   std::vector<CharType> keyword;                                  // relative search: the text to look for
   std::vector<CharType> custom_char_seq = {};                     // optional character sequence replacing ASCII order
   CharType wildcard = '*';                                        // keyword character that matches anything

   std::vector<short> reference_values = {};                       // value scan: the numeric sequence

   int preferred_num_threads = std::thread::hardware_concurrency();   // advisory here: the GPU does the scanning
   int preferred_search_block_size = 524288;                          // BINDING: the skip chains restart at every block
   int preferred_preview_width = 50;                                  // elements per preview window
};

// Unscoped, as in the reference (its engine uses the enumerators unqualified).
enum SearchStep { Initializing, Searching, GeneratingPreviews, Aborting };

template <typename DataType>
class SearchEngine {
public:
   // (percentage, step); see the protocol in the header comment
   using ProgressCallback = std::function<void(int, const SearchStep)>;

   explicit SearchEngine(const SearchConfig &cfg) : config(cfg) {}

   // Streams the file to the GPU in slabs of whole blocks (pinned staging, reads overlapped with copies and scans)
   // and returns the matches in ascending offset order.
   std::vector<SearchResult<DataType>> run(ProgressCallback on_progress, std::atomic<bool> &abort_flag,
                                           bool generate_previews = false);

private:
   SearchConfig config;

   // preview of one match: the window of preferred_preview_width elements centred on it, decoded with its table
   std::string render_preview(std::ifstream &file, uint64_t file_size, uint64_t match_offset,
                              std::map<CharType, DataType> &values_map);
   // '#' for values outside the table, a hex dump for value scans
   std::string render_window(std::map<CharType, DataType> &values_map, std::vector<DataType> &window);
};

}  // namespace mmoore

#endif
