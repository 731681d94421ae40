// search_engine.cpp -- mmoore::SearchEngine<T> over the C-ABI (include/mmoore_b200.h).
//
// Replaces the reference's src/core/search_engine.cpp.  What is gone: one ifstream + one
// std::async thread per block, a full buffer copy per alignment, the in-place byte swap and the
// dispatcher's 5 ms polling sleep (:104-188).  What replaces it: the file is read in slabs of whole
// blocks into two page-locked buffers (several threads pread() disjoint pieces of a slab) and each
// slab is handed to mmg_engine_scan_async, which copies it to the GPU and runs every (block,
// alignment) chain there in one pass; while slab k is copied and scanned the host already reads
// slab k+1, and the library's two internal streams let the copy of k+1 overlap the scan of k.
// Offsets come back already sorted.
// What is kept exactly: block geometry (:218-253), the callback protocol and abort contract
// (:47, :80, :161-165, :177-187, :191), "File not found" (:43-45), result order (:193-197) and the
// preview text (:256-348, including the sticky stream state of the shared ifstream).
#include "mmoore/search_engine.hpp"

#include "mmoore_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

[[noreturn]] void throw_last(int rc) {
   if (rc == MMG_ERR_SKIP_OOB) throw std::runtime_error("Skip table index out of bounds");
   const char *msg = mmg_last_error();
   throw std::runtime_error(msg && *msg ? msg : "monkey-moore GPU search failed");
}

// UTF-8 encoding of one code point (the reference goes through std::codecvt_utf8<char32_t>,
// src/core/encoding.hpp:25-28)
std::string to_utf8(char32_t cp) {
   // std::wstring_convert<std::codecvt_utf8<char32_t>, char32_t>::to_bytes throws std::range_error for code points that
   // have no UTF-8 form: surrogates and everything above U+10FFFF
   if ((cp >= 0xD800 && cp <= 0xDFFF) || cp > 0x10FFFF) throw std::range_error("wstring_convert::to_bytes");
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
   uint64_t size = 0;
   bool pinned = false;
   explicit PinnedBuffer(uint64_t n) : size(n) {
      ptr = static_cast<uint8_t *>(mmg_host_alloc(n));
      pinned = ptr != nullptr;
      if (!ptr) ptr = static_cast<uint8_t *>(::operator new(n ? n : 1));
   }
   ~PinnedBuffer() {
      if (pinned) mmg_host_free(ptr);
      else ::operator delete(ptr);
   }
};

// Page-locking memory costs milliseconds per slab, a GUI session runs search after search: the staging slabs are
// kept in a small process-wide pool instead of being pinned and released by every run().
class StagingPool {
 public:
   std::unique_ptr<PinnedBuffer> take(uint64_t n) {
      {
         std::lock_guard<std::mutex> lock(mutex_);
         for (size_t i = 0; i < free_.size(); i++)
            if (free_[i]->size >= n) {
               auto b = std::move(free_[i]);
               free_.erase(free_.begin() + static_cast<std::ptrdiff_t>(i));
               return b;
            }
      }
      return std::make_unique<PinnedBuffer>(n);
   }
   void give(std::unique_ptr<PinnedBuffer> b) {
      if (!b) return;
      std::lock_guard<std::mutex> lock(mutex_);
      if (free_.size() < 4) free_.push_back(std::move(b));      // at most 4 slabs stay pinned
   }
 private:
   std::mutex mutex_;
   std::vector<std::unique_ptr<PinnedBuffer>> free_;
};

StagingPool &staging_pool() {
   static StagingPool *pool = new StagingPool();      // never destroyed: no CUDA calls during static destruction
   return *pool;
}

// bytes of whole blocks handed to one GPU scan (MMOORE_SLAB_MIB: development aid)
static const uint64_t kSlabBytes = (std::getenv("MMOORE_SLAB_MIB") ? std::max(1, std::atoi(std::getenv("MMOORE_SLAB_MIB"))) : 32) * (1ull << 20);
constexpr uint64_t kReadPiece = 2ull << 20;     // smallest piece one reader thread takes

// Reads file bytes [lo, lo + len) into dst with several threads (pread on one descriptor); bytes the file no longer
// has (it shrank underneath us) are zero.
void read_range(int fd, uint64_t lo, uint64_t len, uint8_t *dst) {
   auto piece = [fd](uint64_t at, uint64_t n, uint8_t *out) {
      uint64_t done = 0;
      while (done < n) {
         const ssize_t got = ::pread(fd, out + done, n - done, static_cast<off_t>(at + done));
         if (got <= 0) break;
         done += static_cast<uint64_t>(got);
      }
      if (done < n) std::memset(out + done, 0, n - done);
   };
   const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
   const uint64_t want = std::min<uint64_t>(std::min<uint64_t>(hw, 16), (len + kReadPiece - 1) / kReadPiece);
   if (want <= 1) { piece(lo, len, dst); return; }
   const uint64_t step = ((len + want - 1) / want + 4095) & ~4095ull;
   std::vector<std::thread> readers;
   for (uint64_t at = step; at < len; at += step)
      readers.emplace_back(piece, lo + at, std::min(step, len - at), dst + at);
   piece(lo, std::min(step, len), dst);
   for (auto &t : readers) t.join();
}

// MMOORE_PROFILE=1: host-side phase times of run() on stderr (development aid; no effect on results)
struct PhaseClock {
   bool on = std::getenv("MMOORE_PROFILE") != nullptr;
   double t[4] = {0, 0, 0, 0};
   double born = now();
   static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
   template <class F> auto time(int k, F &&f) {
      if (!on) return f();
      const double t0 = now();
      struct Stop { double &acc; double t0; ~Stop() { acc += now() - t0; } } stop{t[k], t0};
      return f();
   }
   ~PhaseClock() {
      if (on) std::fprintf(stderr, "[mmoore] run(): take/pin %.2f ms  read %.2f ms  enqueue %.2f ms  collect %.2f ms  (slab loop %.2f ms)\n",
                           1e3 * t[0], 1e3 * t[1], 1e3 * t[2], 1e3 * t[3], 1e3 * (now() - born));
   }
};

struct FileDescriptor {
   int fd;
   explicit FileDescriptor(const std::filesystem::path &p) : fd(::open(p.c_str(), O_RDONLY)) {}
   ~FileDescriptor() { if (fd >= 0) ::close(fd); }
};

// Read-only mapping of the whole file: slabs are then copied into the pinned staging buffers by the library's pool of
// copy threads (mmg_host_copy) -- on the boxes measured, 60+ GB/s from the page cache against 28 GB/s for eight pread()
// threads and 5 GB/s for one.  No mapping (empty file, mmap refused): read_range() and its pread() threads take over.
struct FileMapping {
   const uint8_t *base = nullptr;
   uint64_t size = 0;
   FileMapping(int fd, uint64_t n) {
      if (n == 0) return;
      void *p = ::mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
      if (p == MAP_FAILED) return;
      ::madvise(p, n, MADV_SEQUENTIAL);
      base = static_cast<const uint8_t *>(p);
      size = n;
   }
   // Tearing down the page tables of a large mapping takes milliseconds (512 MiB: ~2 ms) and nobody waits for it:
   // a detached thread does it while run() returns its results.
   ~FileMapping() {
      if (!base) return;
      void *p = const_cast<uint8_t *>(base);
      const uint64_t n = size;
      if (n < (64ull << 20)) { ::munmap(p, n); return; }
      try {
         std::thread([p, n] { ::munmap(p, n); }).detach();
      } catch (...) {
         ::munmap(p, n);
      }
   }
   // the file must still be as long as when it was mapped (touching pages past a truncation would fault)
   bool intact(int fd) const {
      struct stat st;
      return base && ::fstat(fd, &st) == 0 && static_cast<uint64_t>(st.st_size) >= size;
   }
};

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
      FileDescriptor file(config.file_path);
      if (file.fd < 0) throw std::runtime_error("Worker thread failed to open file: " + config.file_path.string());
      FileMapping mapping(file.fd, file_size);

      const uint64_t blocks_per_slab = std::max<uint64_t>(1, kSlabBytes / block);
      const uint64_t slab_capacity = std::min<uint64_t>(file_size, blocks_per_slab * block + overlap);
      const bool big_endian = config.endianness == mmoore::Endianness::Big;

      // two slabs in flight: [read k+1 on the host]  ||  [H2D + scan of k on the GPU]
      struct Slot {
         std::unique_ptr<PinnedBuffer> buf;
         mmg_results *res = nullptr;
         uint64_t blocks = 0;
      } slots[2];
      struct Pending {     // pending scans are completed (or dropped) before the buffers they read go back to the pool
         Slot *s;
         ~Pending() {
            for (int i = 0; i < 2; i++) {
               if (s[i].res) mmg_results_free(s[i].res);
               staging_pool().give(std::move(s[i].buf));
            }
         }
      } pending{slots};

      // completes the scan of a slot: results in file order, then one callback per block of the slab, exactly as the
      // reference's workers report (float accumulation included); false = aborted
      auto collect = [&](Slot &slot) -> bool {
         mmg_results *res = slot.res;
         slot.res = nullptr;
         const int rc = mmg_results_wait(res);
         if (rc != MMG_OK) { mmg_results_free(res); throw_last(rc); }
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
         for (uint64_t b = 0; b < slot.blocks; b++) {
            total_progress += progress_increment;
            on_progress(static_cast<int>(total_progress), SearchStep::Searching);
            if (abort_flag) return false;
         }
         return true;
      };

      PhaseClock clock;
      uint64_t k = 0;
      for (uint64_t first = 0; first < num_blocks; first += blocks_per_slab, k++) {
         Slot &slot = slots[k & 1];
         if (slot.res && !clock.time(3, [&] { return collect(slot); })) return {};          // slab k-2 used this buffer
         if (abort_flag) return {};
         if (!slot.buf) clock.time(0, [&] { slot.buf = staging_pool().take(slab_capacity); return 0; });
         const uint64_t n = std::min<uint64_t>(blocks_per_slab, num_blocks - first);
         const uint64_t lo = first * block;
         const uint64_t hi = std::min<uint64_t>(file_size, (first + n) * static_cast<uint64_t>(block) + overlap);
         clock.time(1, [&] {
            if (mapping.intact(file.fd)) mmg_host_copy(slot.buf->ptr, mapping.base + lo, hi - lo);
            else read_range(file.fd, lo, hi - lo, slot.buf->ptr);
            return 0;
         });
         const int rc = clock.time(2, [&] {
            return mmg_engine_scan_async(searcher->program(), slot.buf->ptr, hi - lo, MMG_MEM_HOST, file_size, block, first, n,
                                         big_endian ? 1 : 0, &slot.res);
         });
         if (rc != MMG_OK) throw_last(rc);
         slot.blocks = n;
      }
      // the last two slabs, oldest first
      for (uint64_t j = k >= 2 ? k - 2 : 0; j < k; j++)
         if (slots[j & 1].res && !clock.time(3, [&] { return collect(slots[j & 1]); })) return {};
   }
   if (abort_flag) return {};

   on_progress(100, GeneratingPreviews);
   // slabs arrive in file order and each scan returns ascending offsets: already sorted

   if (generate_previews && !results.empty()) {
      std::ifstream preview_file(config.file_path, std::ios::binary);
      if (!preview_file.is_open())
         throw std::runtime_error("Failed to open file to generate previews: " + config.file_path.string());
      for (auto &result : results) result.preview = render_preview(preview_file, file_size, result.offset, result.values_map);
   }
   return results;
}

template <typename DataType>
std::string mmoore::SearchEngine<DataType>::render_preview(std::ifstream &file, uint64_t file_size, uint64_t match_offset,
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
   return render_window(values_map, window);
}

template <typename DataType>
std::string mmoore::SearchEngine<DataType>::render_window(std::map<CharType, DataType> &values_map,
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
