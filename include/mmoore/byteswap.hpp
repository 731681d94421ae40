// byteswap.hpp -- endianness vocabulary of the search API.
//
// Same public surface as the reference's include/mmoore/byteswap.hpp:11-79 (Endianness,
// get_system_endianness, swap_always, swap_on_little_endian, swap_on_big_endian,
// adjust_endianness).  On the GPU path the swap itself is a byte permute inside the filter
// kernel's load (csrc/scan_kernels.cu, extract<>); these host helpers remain for callers
// (GUI, previews, tests) that use them directly.
#ifndef MMOORE_B200_BYTESWAP_HPP
#define MMOORE_B200_BYTESWAP_HPP

#include <cstddef>
#include <cstdint>

namespace mmoore {

enum class Endianness { Little, Big };

inline Endianness get_system_endianness() {
   const uint16_t probe = 0x0102;
   return *reinterpret_cast<const uint8_t *>(&probe) == 0x02 ? Endianness::Little : Endianness::Big;
}

// Identity for single-byte types; byte reversal for 16- and 32-bit values.
template <typename T>
constexpr T swap_always(T value) {
   if constexpr (sizeof(T) == 2) {
      return static_cast<T>(((value & 0x00FFu) << 8) | ((value >> 8) & 0x00FFu));
   } else if constexpr (sizeof(T) == 4) {
      return static_cast<T>(((value & 0x000000FFu) << 24) | ((value & 0x0000FF00u) << 8) |
                            ((value >> 8) & 0x0000FF00u) | ((value >> 24) & 0x000000FFu));
   } else {
      return value;
   }
}

template <typename T>
T swap_on_little_endian(T value) {
   return get_system_endianness() == Endianness::Little ? swap_always(value) : value;
}

template <typename T>
T swap_on_big_endian(T value) {
   return get_system_endianness() == Endianness::Big ? swap_always(value) : value;
}

// In-place conversion of `count` elements to host order from `desired_endianness` (or back).
template <typename T>
void adjust_endianness(T *dataPtr, size_t count, Endianness desired_endianness) {
   if (get_system_endianness() == desired_endianness) return;
   for (size_t i = 0; i < count; ++i) dataPtr[i] = swap_always<T>(dataPtr[i]);
}

}  // namespace mmoore

#endif
