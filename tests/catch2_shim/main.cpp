// Runner of the minimal Catch2 stand-in: executes every registered TEST_CASE once per leaf path of
// its SECTION / GENERATE tree.  Usage: ref_unit_tests [substring-filter]
#include <catch2/catch_test_macros.hpp>

#include <cstring>

namespace catch_shim {

State &state() {
   static State s;
   return s;
}

std::vector<TestCase> &registry() {
   static std::vector<TestCase> r;
   return r;
}

void report_failure(const char *kind, const char *expr, const char *file, int line, const std::string &extra) {
   State &s = state();
   s.failures++;
   std::cout << file << ":" << line << ": FAILED: " << kind << "( " << expr << " )\n  in test case: " << s.current_test << "\n";
   for (const auto &n : s.section_names) std::cout << "    section: " << n << "\n";
   if (!s.info.empty()) std::cout << "    info: " << s.info << "\n";
   if (!extra.empty()) std::cout << "    " << extra << "\n";
}

}  // namespace catch_shim

int main(int argc, char **argv) {
   using namespace catch_shim;
   const char *filter = argc > 1 ? argv[1] : nullptr;
   int cases = 0, runs = 0, failed_cases = 0;
   for (const TestCase &tc : registry()) {
      if (filter && !std::strstr(tc.name, filter)) continue;
      cases++;
      State &s = state();
      s.current_test = tc.name;
      s.target.clear();
      s.sib.clear();
      const int failures_before = s.failures;
      for (;;) {
         s.depth = 0;
         s.count.assign(1, 0);
         s.info.clear();
         s.section_names.clear();
         runs++;
         try {
            tc.fn();
         } catch (const AbortRun &) {
         } catch (const std::exception &e) {
            report_failure("unexpected exception", e.what(), "-", 0);
         } catch (...) {
            report_failure("unexpected exception", "unknown", "-", 0);
         }
         for (size_t L = 0; L < s.count.size(); L++) {      // levels still open at the end of the run
            if (s.sib.size() <= L) s.sib.resize(L + 1, 0);
            s.sib[L] = s.count[L];
         }
         // next leaf path: advance the deepest level that still has siblings
         while (!s.target.empty()) {
            const size_t d = s.target.size() - 1;
            if (d < s.sib.size() && s.target[d] + 1 < s.sib[d]) { s.target[d]++; break; }
            s.target.pop_back();
         }
         if (s.target.empty()) break;
      }
      if (s.failures != failures_before) failed_cases++;
      std::cout << (s.failures != failures_before ? "[FAILED] " : "[  ok  ] ") << tc.name << "\n";
   }
   State &s = state();
   std::cout << "test cases: " << cases << " | failed: " << failed_cases << " | runs: " << runs
             << " | assertions: " << s.assertions << " | failures: " << s.failures << "\n";
   return s.failures == 0 && cases > 0 ? 0 : 1;
}
