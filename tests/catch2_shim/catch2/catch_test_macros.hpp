// Minimal stand-in for Catch2 v3's <catch2/catch_test_macros.hpp> (Catch2 is not installed in this
// image).  It implements exactly what the reference's tests use -- TEST_CASE, nested SECTION with
// one leaf path per run, GENERATE (cartesian re-runs), REQUIRE / CHECK / REQUIRE_THAT /
// REQUIRE_THROWS_AS / FAIL / INFO / CAPTURE -- so that /root/reference/tests/*.cpp compile
// UNMODIFIED against this repo's include/mmoore headers and run against the GPU library.
// TEST INFRASTRUCTURE ONLY.
#ifndef CATCH2_SHIM_TEST_MACROS_HPP
#define CATCH2_SHIM_TEST_MACROS_HPP

#include <exception>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace catch_shim {

struct AbortRun {};   // thrown by a failing REQUIRE / FAIL

struct State {
   std::vector<int> target;   // index chosen at each nesting level for the current run
   std::vector<int> count;    // siblings seen at each open level during the current run
   std::vector<int> sib;      // siblings at each level along the path this run took (survives scope exits)
   int depth = 0;
   int failures = 0;
   int assertions = 0;
   std::string current_test;
   std::string info;
   std::vector<std::string> section_names;
};

State &state();

struct TestCase {
   const char *name;
   void (*fn)();
};
std::vector<TestCase> &registry();

struct Registrar {
   Registrar(const char *name, void (*fn)()) { registry().push_back({name, fn}); }
};

// A SECTION or one value of a GENERATE is a child at the current nesting level; a run enters
// exactly the child selected by target[level] (default: the first).
struct SectionGuard {
   bool entered = false;
   int saved_depth = 0;
   explicit SectionGuard(const char *name) {
      State &s = state();
      const int d = s.depth;
      if ((int)s.count.size() <= d) s.count.resize(d + 1, 0);
      const int idx = s.count[d]++;
      if ((int)s.target.size() <= d) s.target.push_back(0);
      if (idx == s.target[d]) {
         entered = true;
         saved_depth = d;
         s.depth = d + 1;
         s.count.resize(d + 2);
         s.count[d + 1] = 0;
         s.section_names.push_back(name);
      }
   }
   ~SectionGuard() {
      if (entered) {
         State &s = state();
         s.depth = saved_depth;
         for (size_t L = saved_depth + 1; L < s.count.size(); L++) {   // remember what this path saw
            if (s.sib.size() <= L) s.sib.resize(L + 1, 0);
            s.sib[L] = s.count[L];
         }
         s.count.resize(saved_depth + 1);   // generator levels opened inside this section end with it
         if (!s.section_names.empty()) s.section_names.pop_back();
      }
   }
   explicit operator bool() const { return entered; }
};

template <class T>
T generate(std::initializer_list<T> values) {
   State &s = state();
   const int d = s.depth;
   if ((int)s.count.size() <= d) s.count.resize(d + 1, 0);
   s.count[d] = (int)values.size();
   if ((int)s.target.size() <= d) s.target.push_back(0);
   const int idx = s.target[d];
   s.depth = d + 1;
   s.count.resize(d + 2);
   s.count[d + 1] = 0;
   return *(values.begin() + idx);
}

void report_failure(const char *kind, const char *expr, const char *file, int line, const std::string &extra = "");

inline void check(bool ok, bool fatal, const char *kind, const char *expr, const char *file, int line) {
   state().assertions++;
   if (ok) return;
   report_failure(kind, expr, file, line);
   if (fatal) throw AbortRun{};
}

}  // namespace catch_shim

#define CATCH_SHIM_CAT2(a, b) a##b
#define CATCH_SHIM_CAT(a, b) CATCH_SHIM_CAT2(a, b)

#define TEST_CASE(...)                                                                                   \
   static void CATCH_SHIM_CAT(catch_shim_test_, __LINE__)();                                              \
   static ::catch_shim::Registrar CATCH_SHIM_CAT(catch_shim_reg_, __LINE__)(                              \
      ::catch_shim::first_arg(__VA_ARGS__), &CATCH_SHIM_CAT(catch_shim_test_, __LINE__));                 \
   static void CATCH_SHIM_CAT(catch_shim_test_, __LINE__)()

namespace catch_shim {
inline const char *first_arg(const char *name) { return name; }
inline const char *first_arg(const char *name, const char *) { return name; }
}  // namespace catch_shim

#define SECTION(name) if (::catch_shim::SectionGuard CATCH_SHIM_CAT(catch_shim_sec_, __LINE__){name})

#define REQUIRE(...) ::catch_shim::check(static_cast<bool>(__VA_ARGS__), true, "REQUIRE", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK(...) ::catch_shim::check(static_cast<bool>(__VA_ARGS__), false, "CHECK", #__VA_ARGS__, __FILE__, __LINE__)
#define FAIL(msg)                                                                                        \
   do {                                                                                                  \
      std::ostringstream catch_shim_os;                                                                  \
      catch_shim_os << msg;                                                                              \
      ::catch_shim::state().assertions++;                                                                \
      ::catch_shim::report_failure("FAIL", catch_shim_os.str().c_str(), __FILE__, __LINE__);              \
      throw ::catch_shim::AbortRun{};                                                                    \
   } while (0)
#define INFO(msg)                                                                                        \
   do {                                                                                                  \
      std::ostringstream catch_shim_os;                                                                  \
      catch_shim_os << msg;                                                                              \
      ::catch_shim::state().info = catch_shim_os.str();                                                  \
   } while (0)
#define CAPTURE(...) (void)0
#define REQUIRE_THROWS_AS(expr, type)                                                                    \
   do {                                                                                                  \
      bool catch_shim_thrown = false;                                                                    \
      try { (void)(expr); } catch (const type &) { catch_shim_thrown = true; } catch (...) {}            \
      ::catch_shim::check(catch_shim_thrown, true, "REQUIRE_THROWS_AS", #expr ", " #type, __FILE__, __LINE__); \
   } while (0)
#define REQUIRE_THAT(value, matcher)                                                                     \
   ::catch_shim::check((matcher).match(value), true, "REQUIRE_THAT", #value ", " #matcher, __FILE__, __LINE__)

#endif
