// Stand-in for <catch2/matchers/catch_matchers_vector.hpp>: Catch::Matchers::Equals for std::vector.
#ifndef CATCH2_SHIM_MATCHERS_VECTOR_HPP
#define CATCH2_SHIM_MATCHERS_VECTOR_HPP

#include <vector>

namespace Catch {
namespace Matchers {

template <class T>
struct VectorEquals {
   const std::vector<T> &expected;
   bool match(const std::vector<T> &actual) const {
      if (actual.size() != expected.size()) return false;
      for (size_t i = 0; i < actual.size(); i++)
         if (!(actual[i] == expected[i])) return false;
      return true;
   }
};

template <class T>
VectorEquals<T> Equals(const std::vector<T> &expected) {
   return VectorEquals<T>{expected};
}

}  // namespace Matchers
}  // namespace Catch

#endif
