// bench_search.cpp -- the reference's benchmark harness (benchmarks/bench_search.cpp under /root/reference/:
// single in-memory MonkeyMoore<T>::search per iteration, keyword "abcde" and three wildcard placements, 8/16-bit,
// 128 KiB .. 16 MiB x4, std::mt19937(42) uniform data, :11-105) re-stated with std::chrono because Google Benchmark
// is not installed.  It uses ONLY the public API of include/mmoore/*.hpp, so the same file builds twice:
//   benchmarks/bench_search        against this repo's libmonkey-core.so (GPU)
//   oracle/_ref/bench_search_ref   against the unmodified reference sources (CPU baseline, built by oracle/Makefile)
// Extra rows (not in the reference harness): the BASELINE configs' own patterns and SearchEngine::run over a file.
#include "mmoore/monkey_moore.hpp"
#include "mmoore/search_engine.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <random>
#include <string>
#include <thread>
#include <vector>

template <typename DataType>
static std::vector<DataType> generate_data(size_t size_in_bytes) {      // benchmarks/bench_search.cpp:11-22
   std::vector<DataType> data(size_in_bytes / sizeof(DataType));
   std::mt19937 rng(42);
   std::uniform_int_distribution<unsigned int> dist(0, std::numeric_limits<DataType>::max());
   for (auto &v : data) v = static_cast<DataType>(dist(rng));
   return data;
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename DataType>
static void run_search(const char *name, const std::vector<CharType> &keyword, CharType wildcard, size_t bytes, double budget) {
   auto data = generate_data<DataType>(bytes);
   MonkeyMoore<DataType> searcher(keyword, wildcard, {});
   size_t matches = 0;
   for (int i = 0; i < 2; i++) matches = searcher.search(data.data(), data.size()).size();     // warm-up
   int iters = 0;
   double best = 1e30, total = 0;
   const double t_end = now() + budget;
   do {
      const double t0 = now();
      auto results = searcher.search(data.data(), data.size());
      const double dt = now() - t0;
      matches = results.size();
      best = dt < best ? dt : best;
      total += dt;
      iters++;
   } while (now() < t_end && iters < 200);
   std::printf("%-44s %9zu B  iters %4d  mean %10.1f us  best %10.1f us  %8.3f GB/s  matches %zu\n", name, bytes, iters,
               1e6 * total / iters, 1e6 * best, bytes / (total / iters) / 1e9, matches);
   std::fflush(stdout);
}

template <typename DataType>
static void run_engine(const char *name, const std::vector<CharType> &keyword, CharType wildcard, size_t bytes, int block,
                       const std::string &dir) {
   const std::string path = dir + "/mmoore_bench_engine.bin";
   {
      auto data = generate_data<uint8_t>(bytes);
      std::ofstream f(path, std::ios::binary);
      f.write(reinterpret_cast<const char *>(data.data()), static_cast<std::streamsize>(data.size()));
   }
   mmoore::SearchConfig config;
   config.file_path = path;
   config.keyword = keyword;
   config.wildcard = wildcard;
   config.is_relative_search = true;
   config.preferred_search_block_size = block;
   config.preferred_num_threads = static_cast<int>(std::thread::hardware_concurrency());
   std::atomic<bool> abort_flag{false};
   double best = 1e30;
   size_t matches = 0;
   for (int i = 0; i < 4; i++) {
      mmoore::SearchEngine<DataType> engine(config);
      const double t0 = now();
      auto results = engine.run([](int, const mmoore::SearchStep) {}, abort_flag, false);
      const double dt = now() - t0;
      matches = results.size();
      if (i > 0 && dt < best) best = dt;
   }
   std::remove(path.c_str());
   std::printf("%-44s %9zu B  block %8d  best %10.1f us  %8.3f GB/s  matches %zu  threads %d\n", name, bytes, block, 1e6 * best,
               bytes / best / 1e9, matches, config.preferred_num_threads);
   std::fflush(stdout);
}

int main(int argc, char **argv) {
   const double budget = argc > 1 ? std::atof(argv[1]) : 0.5;      // seconds per row
   const bool engine_only = argc > 2 && std::string(argv[2]) == "engine-only";      // development: only the SearchEngine::run rows
   const bool with_engine = engine_only || (argc > 2 && std::string(argv[2]) == "engine");
   const std::vector<CharType> abcde = {'a', 'b', 'c', 'd', 'e'};
   const std::vector<CharType> front = {'*', 'b', 'c', 'd', 'e'}, middle = {'a', 'b', '*', 'd', 'e'}, end = {'a', 'b', 'c', 'd', '*'};
   // ->RangeMultiplier(4)->Range(128<<10, 16<<20): Google Benchmark appends the upper bound, so the rows are
   // 128 KiB, 512 KiB, 2 MiB, 8 MiB and 16 MiB (benchmarks/bench_search.cpp:67-105 of the reference)
   const std::vector<CharType> monkey = {'m', 'o', 'n', 'k', 'e', 'y'}, mokeys = {'m', 'o', '*', 'k', 'e', 'y', '*', 's'};
   const size_t sizes[] = {128u << 10, 512u << 10, 2u << 20, 8u << 20, 16u << 20};
   if (!engine_only) {
   for (size_t bytes : sizes) run_search<uint8_t>("BM_Search/Relative/8-Bit", abcde, 0, bytes, budget);
   for (size_t bytes : sizes) run_search<uint16_t>("BM_Search/Relative/16-Bit", abcde, 0, bytes, budget);
   for (size_t bytes : sizes) run_search<uint8_t>("BM_Search/Relative/Wildcard/Front/8-Bit", front, '*', bytes, budget);
   for (size_t bytes : sizes) run_search<uint8_t>("BM_Search/Relative/Wildcard/Middle/8-Bit", middle, '*', bytes, budget);
   for (size_t bytes : sizes) run_search<uint8_t>("BM_Search/Relative/Wildcard/Back/8-Bit", end, '*', bytes, budget);
   for (size_t bytes : sizes) run_search<uint16_t>("BM_Search/Relative/Wildcard/Front/16-Bit", front, '*', bytes, budget);
   for (size_t bytes : sizes) run_search<uint16_t>("BM_Search/Relative/Wildcard/Middle/16-Bit", middle, '*', bytes, budget);
   for (size_t bytes : sizes) run_search<uint16_t>("BM_Search/Relative/Wildcard/Back/16-Bit", end, '*', bytes, budget);
   // the BASELINE configs' own patterns (not part of the reference harness)
   run_search<uint8_t>("cfg1 8-bit monkey", monkey, 0, 16u << 20, budget);
   run_search<uint16_t>("cfg2 16-bit mo*key*s", mokeys, '*', 16u << 20, budget);
   }
   if (with_engine) {
      const char *shm = "/dev/shm";
      run_engine<uint8_t>("SearchEngine::run 8-bit monkey", monkey, 0, 512u << 20, 524288, shm);
      run_engine<uint8_t>("SearchEngine::run 8-bit monkey", monkey, 0, 512u << 20, 8388608, shm);
      run_engine<uint16_t>("SearchEngine::run 16-bit mo*key*s", mokeys, '*', 512u << 20, 524288, shm);
      run_engine<uint16_t>("SearchEngine::run 16-bit mo*key*s", mokeys, '*', 512u << 20, 8388608, shm);
   }
   return 0;
}
