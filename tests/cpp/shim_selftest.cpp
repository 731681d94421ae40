#include <catch2/catch_test_macros.hpp>
#include <catch2/generators/catch_generators.hpp>
#include <set>
static std::multiset<std::string> seen;
TEST_CASE("shim: nested sections and generators", "[x]") {
   std::string path = "r";
   SECTION("a") { path += "a";
      SECTION("a1") { path += "1"; }
      SECTION("a2") { path += "2"; int g = GENERATE(7, 8, 9); int h = GENERATE(1, 2); path += std::to_string(g) + std::to_string(h); }
   }
   SECTION("b") { path += "b"; }
   SECTION("c") { path += "c"; SECTION("c1") { path += "1"; SECTION("deep") { path += "d"; } } }
   seen.insert(path);
}
TEST_CASE("shim: verify", "[x]") {
   std::multiset<std::string> expect = {"ra1","ra271","ra272","ra281","ra282","ra291","ra292","rb","rc1d"};
   CHECK(seen == expect);
   for (auto &s : seen) std::cout << s << " "; std::cout << "\n";
}
