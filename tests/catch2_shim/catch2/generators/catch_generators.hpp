// Stand-in for <catch2/generators/catch_generators.hpp>: GENERATE(v1, v2, ...) re-runs the test case
// once per value (cartesian with every other GENERATE / SECTION choice).
#ifndef CATCH2_SHIM_GENERATORS_HPP
#define CATCH2_SHIM_GENERATORS_HPP

#include "../catch_test_macros.hpp"

#define GENERATE(...) ::catch_shim::generate({__VA_ARGS__})

#endif
