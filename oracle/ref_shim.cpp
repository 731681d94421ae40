// ref_shim.cpp -- C-ABI doorway into the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
// /root/reference/src/core/{monkey_moore,search_engine}.cpp (in place, never
// copied) into oracle/_ref/libmmref.so.  It is used to (a) validate the C
// restatement in mm_oracle.c, (b) generate the golden fixtures under
// tests/golden/, and (c) as the CPU baseline of bench.py (--impl reference,
// cpu_baseline.kind == "reference").  The product never loads it.
#include "mmoore/monkey_moore.hpp"
#include "mmoore/search_engine.hpp"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

template <class Ty>
struct SearchOut {
   std::vector<uint64_t> pos;
   std::vector<uint32_t> keys;     // flattened maps: per match `table_size` (key,value) pairs
   std::vector<uint32_t> values;
   std::vector<uint32_t> sizes;
};

template <class Ty>
MonkeyMoore<Ty> *make_searcher(const uint32_t *kw, int L, uint32_t wildcard, const uint32_t *seq, int nseq,
                               const int16_t *vals, int nvals) {
   if (vals != nullptr) {
      std::vector<short> v(vals, vals + nvals);
      return new MonkeyMoore<Ty>(v);
   }
   std::vector<CharType> keyword(kw, kw + L);
   std::vector<CharType> char_seq(seq, seq + nseq);
   return new MonkeyMoore<Ty>(keyword, static_cast<CharType>(wildcard), char_seq);
}

template <class Ty>
int64_t do_search(const uint32_t *kw, int L, uint32_t wildcard, const uint32_t *seq, int nseq,
                  const int16_t *vals, int nvals, const void *data, uint64_t n,
                  uint64_t *out_pos, uint32_t *out_sizes, uint32_t *out_keys, uint32_t *out_values,
                  uint64_t cap_matches, uint64_t cap_entries) {
   std::unique_ptr<MonkeyMoore<Ty>> s(make_searcher<Ty>(kw, L, wildcard, seq, nseq, vals, nvals));
   auto results = s->search(static_cast<const Ty *>(data), n);
   uint64_t e = 0;
   for (size_t i = 0; i < results.size(); i++) {
      if (i < cap_matches) {
         if (out_pos) out_pos[i] = results[i].first;
         if (out_sizes) out_sizes[i] = static_cast<uint32_t>(results[i].second.size());
      }
      for (const auto &kv : results[i].second) {
         if (e < cap_entries) {
            if (out_keys) out_keys[e] = static_cast<uint32_t>(kv.first);
            if (out_values) out_values[e] = static_cast<uint32_t>(kv.second);
         }
         e++;
      }
   }
   return static_cast<int64_t>(results.size());
}

template <class Ty>
double time_search(const uint32_t *kw, int L, uint32_t wildcard, const uint32_t *seq, int nseq,
                   const int16_t *vals, int nvals, const void *data, uint64_t n, int iters, int64_t *matches) {
   std::unique_ptr<MonkeyMoore<Ty>> s(make_searcher<Ty>(kw, L, wildcard, seq, nseq, vals, nvals));
   double best = 1e300;
   int64_t m = 0;
   for (int it = 0; it < iters; it++) {
      auto t0 = std::chrono::steady_clock::now();
      auto r = s->search(static_cast<const Ty *>(data), n);
      auto t1 = std::chrono::steady_clock::now();
      double dt = std::chrono::duration<double>(t1 - t0).count();
      if (dt < best) best = dt;
      m = static_cast<int64_t>(r.size());
   }
   if (matches) *matches = m;
   return best;
}

struct EngineOut {
   std::vector<uint64_t> offsets;
   std::vector<uint32_t> sizes, keys, values;
   std::vector<std::string> previews;
   std::vector<int> progress_pct, progress_step;
};

template <class Ty>
EngineOut *do_engine(const char *path, int is_relative, int big_endian, const uint32_t *kw, int L,
                     uint32_t wildcard, const uint32_t *seq, int nseq, const int16_t *vals, int nvals,
                     int threads, int block, int preview_width, int previews, int abort_after) {
   mmoore::SearchConfig cfg;
   cfg.file_path = path;
   cfg.is_relative_search = is_relative != 0;
   cfg.endianness = big_endian ? mmoore::Endianness::Big : mmoore::Endianness::Little;
   cfg.keyword.assign(kw, kw + L);
   cfg.custom_char_seq.assign(seq, seq + nseq);
   cfg.wildcard = static_cast<CharType>(wildcard);
   cfg.reference_values.assign(vals, vals + nvals);
   cfg.preferred_num_threads = threads;
   cfg.preferred_search_block_size = block;
   cfg.preferred_preview_width = preview_width;

   auto *out = new EngineOut();
   std::atomic<bool> abort_flag{false};
   int calls = 0;
   mmoore::SearchEngine<Ty> engine(cfg);
   auto results = engine.run(
      [&](int pct, const mmoore::SearchStep step) {
         // the reference invokes this under its own progress mutex or from the calling thread
         out->progress_pct.push_back(pct);
         out->progress_step.push_back(static_cast<int>(step));
         calls++;
         if (abort_after > 0 && calls >= abort_after) abort_flag = true;
      },
      abort_flag, previews != 0);
   for (auto &r : results) {
      out->offsets.push_back(r.offset);
      out->sizes.push_back(static_cast<uint32_t>(r.values_map.size()));
      for (const auto &kv : r.values_map) {
         out->keys.push_back(static_cast<uint32_t>(kv.first));
         out->values.push_back(static_cast<uint32_t>(kv.second));
      }
      out->previews.push_back(r.preview);
   }
   return out;
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_error.c_str(); }

// Returns number of matches, or -1 when the reference threw (message in ref_last_error()).
int64_t ref_search(int bits, const uint32_t *kw, int L, uint32_t wildcard, const uint32_t *seq, int nseq,
                   const int16_t *vals, int nvals, const void *data, uint64_t n,
                   uint64_t *out_pos, uint32_t *out_sizes, uint32_t *out_keys, uint32_t *out_values,
                   uint64_t cap_matches, uint64_t cap_entries) {
   try {
      if (bits == 8)
         return do_search<uint8_t>(kw, L, wildcard, seq, nseq, vals, nvals, data, n, out_pos, out_sizes,
                                   out_keys, out_values, cap_matches, cap_entries);
      return do_search<uint16_t>(kw, L, wildcard, seq, nseq, vals, nvals, data, n, out_pos, out_sizes,
                                 out_keys, out_values, cap_matches, cap_entries);
   } catch (const std::exception &e) {
      g_error = e.what();
      return -1;
   }
}

// Constructs the searcher only; 0 on success, -1 if the reference threw.
int ref_compile(int bits, const uint32_t *kw, int L, uint32_t wildcard, const uint32_t *seq, int nseq,
                const int16_t *vals, int nvals) {
   try {
      if (bits == 8) delete make_searcher<uint8_t>(kw, L, wildcard, seq, nseq, vals, nvals);
      else delete make_searcher<uint16_t>(kw, L, wildcard, seq, nseq, vals, nvals);
      return 0;
   } catch (const std::exception &e) {
      g_error = e.what();
      return -1;
   }
}

// Times `iters` calls of MonkeyMoore<Ty>::search on one in-memory buffer (the
// re-statement of benchmarks/bench_search.cpp:24-38 without Google Benchmark).
// Returns best seconds per call; *matches receives the match count.
double ref_time_search(int bits, const uint32_t *kw, int L, uint32_t wildcard, const uint32_t *seq, int nseq,
                       const int16_t *vals, int nvals, const void *data, uint64_t n, int iters,
                       int64_t *matches) {
   try {
      if (bits == 8) return time_search<uint8_t>(kw, L, wildcard, seq, nseq, vals, nvals, data, n, iters, matches);
      return time_search<uint16_t>(kw, L, wildcard, seq, nseq, vals, nvals, data, n, iters, matches);
   } catch (const std::exception &e) {
      g_error = e.what();
      return -1.0;
   }
}

// Runs mmoore::SearchEngine<T>::run on a file.  Returns an opaque handle (NULL if the reference threw).
void *ref_engine_run(int bits, const char *path, int is_relative, int big_endian, const uint32_t *kw, int L,
                     uint32_t wildcard, const uint32_t *seq, int nseq, const int16_t *vals, int nvals,
                     int threads, int block, int preview_width, int previews, int abort_after) {
   try {
      if (bits == 8)
         return do_engine<uint8_t>(path, is_relative, big_endian, kw, L, wildcard, seq, nseq, vals, nvals,
                                   threads, block, preview_width, previews, abort_after);
      return do_engine<uint16_t>(path, is_relative, big_endian, kw, L, wildcard, seq, nseq, vals, nvals,
                                 threads, block, preview_width, previews, abort_after);
   } catch (const std::exception &e) {
      g_error = e.what();
      return nullptr;
   }
}

uint64_t ref_engine_count(void *h) { return static_cast<EngineOut *>(h)->offsets.size(); }
uint64_t ref_engine_entries(void *h) { return static_cast<EngineOut *>(h)->keys.size(); }
uint64_t ref_engine_progress_count(void *h) { return static_cast<EngineOut *>(h)->progress_pct.size(); }

void ref_engine_get(void *h, uint64_t *offsets, uint32_t *sizes, uint32_t *keys, uint32_t *values,
                    int *progress_pct, int *progress_step) {
   auto *o = static_cast<EngineOut *>(h);
   if (offsets) std::memcpy(offsets, o->offsets.data(), o->offsets.size() * sizeof(uint64_t));
   if (sizes) std::memcpy(sizes, o->sizes.data(), o->sizes.size() * sizeof(uint32_t));
   if (keys) std::memcpy(keys, o->keys.data(), o->keys.size() * sizeof(uint32_t));
   if (values) std::memcpy(values, o->values.data(), o->values.size() * sizeof(uint32_t));
   if (progress_pct) std::memcpy(progress_pct, o->progress_pct.data(), o->progress_pct.size() * sizeof(int));
   if (progress_step) std::memcpy(progress_step, o->progress_step.data(), o->progress_step.size() * sizeof(int));
}

const char *ref_engine_preview(void *h, uint64_t i) { return static_cast<EngineOut *>(h)->previews[i].c_str(); }

void ref_engine_free(void *h) { delete static_cast<EngineOut *>(h); }

}  // extern "C"
