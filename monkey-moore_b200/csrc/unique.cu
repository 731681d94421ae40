// unique.cu -- distinct inferred tables of a match list, on the device.
//
// The reference's GUI shows, by default, only the first result of every distinct character table: it walks
// the sorted result vector and keeps a result iff its equivalency_map is not yet in a `unique` vector
// (/root/reference/src/gui/monkey_frame.cpp:1236-1245) -- O(results x distinct tables) map comparisons on the host.
// Here the table of a match is a function of the one or two raw element values the scan already emitted per match
// (the u32 "values" word v0 | v1 << 16, see mmg_program_table), so "first match of every distinct table" is a
// min-reduction of the match index keyed by (a masked form of) that word:
//   * key space <= 65536 (8-bit searches, and every search whose table depends on v0 only): a direct-address table
//     of u64 first indices; a thread reads the slot first and issues the atomicMin only when it would lower it, so
//     a list of 10^9 matches over a few hundred tables costs one streaming read of the values, not 10^9
//     same-address atomics;
//   * 16-bit mixed-case searches (v0 and v1 both matter): open addressing over u64 slots (key << 32 | index).
#include "launch.h"

#include <algorithm>
#include <cstdint>

namespace {

#define UQ_EMPTY 0xFFFFFFFFFFFFFFFFull

__global__ void __launch_bounds__(256) k_unique_direct(const uint32_t *val, uint64_t n, uint32_t keymask, uint32_t pack8,
                                                       unsigned long long *first) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t k = val[i] & keymask;
        if (pack8) k = (k & 0xFFu) | ((k >> 8) & 0xFF00u);      // 8-bit pairs: v0 | v1 << 8
        k &= 0xFFFFu;
        if (first[k] > i) atomicMin(first + k, (unsigned long long)i);
    }
}

__device__ __forceinline__ uint32_t uq_hash(uint32_t k) {
    k ^= k >> 16; k *= 0x7FEB352Du; k ^= k >> 15; k *= 0x846CA68Bu; k ^= k >> 16;
    return k;
}

// slots: key << 32 | first index (indices < 2^32 - 1); capacity is a power of two, at least twice the distinct keys
__global__ void __launch_bounds__(256) k_unique_hash(const uint32_t *val, uint64_t n, unsigned long long *slots,
                                                     uint32_t capmask, unsigned int *overflow) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = val[i];
        const unsigned long long mine = ((unsigned long long)k << 32) | (unsigned long long)i;
        uint32_t h = uq_hash(k) & capmask;
        for (uint32_t probes = 0;; probes++) {
            unsigned long long cur = slots[h];
            if (cur == UQ_EMPTY) {
                cur = atomicCAS(slots + h, UQ_EMPTY, mine);
                if (cur == UQ_EMPTY) break;
            }
            if ((uint32_t)(cur >> 32) == k) {
                if (cur > mine) atomicMin(slots + h, mine);
                break;
            }
            if (probes > capmask) { atomicExch(overflow, 1u); break; }
            h = (h + 1) & capmask;
        }
    }
}

// appends the first indices of all occupied slots to out (unordered; the host sorts the usually short list)
__global__ void __launch_bounds__(256) k_unique_collect(const unsigned long long *slots, uint64_t nslots, uint32_t low32,
                                                        unsigned long long *out, unsigned long long *count, uint64_t capacity) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += stride) {
        const unsigned long long s = slots[i];
        if (s == UQ_EMPTY) continue;
        const unsigned long long at = atomicAdd(count, 1ull);
        if (at < capacity) out[at] = low32 ? (s & 0xFFFFFFFFull) : s;
    }
}

}  // namespace

cudaError_t mmg_launch_unique_direct(const uint32_t *val, uint64_t n, uint32_t keymask, bool pack8, uint64_t *first,
                                     cudaStream_t stream) {
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148ull * 8);
    k_unique_direct<<<grid, 256, 0, stream>>>(val, n, keymask, pack8 ? 1u : 0u, reinterpret_cast<unsigned long long *>(first));
    return cudaGetLastError();
}

cudaError_t mmg_launch_unique_hash(const uint32_t *val, uint64_t n, uint64_t *slots, uint32_t capmask, unsigned int *overflow,
                                   cudaStream_t stream) {
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148ull * 8);
    k_unique_hash<<<grid, 256, 0, stream>>>(val, n, reinterpret_cast<unsigned long long *>(slots), capmask, overflow);
    return cudaGetLastError();
}

cudaError_t mmg_launch_unique_collect(const uint64_t *slots, uint64_t nslots, bool low32, uint64_t *out, uint64_t *count,
                                      uint64_t capacity, cudaStream_t stream) {
    const unsigned grid = (unsigned)std::min<uint64_t>((nslots + 255) / 256, 148ull * 8);
    k_unique_collect<<<grid, 256, 0, stream>>>(reinterpret_cast<const unsigned long long *>(slots), nslots, low32 ? 1u : 0u,
                                               reinterpret_cast<unsigned long long *>(out),
                                               reinterpret_cast<unsigned long long *>(count), capacity);
    return cudaGetLastError();
}
