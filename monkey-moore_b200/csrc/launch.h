// launch.h -- host-callable launch helpers implemented in scan_kernels.cu
#ifndef MMG_LAUNCH_H
#define MMG_LAUNCH_H

#include "scan_kernels.cuh"

bool mmg_filter_supported(int W, int lag_bytes);
cudaError_t mmg_filter_occupancy(int W, int lag_bytes, bool be, int nkeys, int *blocks_per_sm);
cudaError_t mmg_launch_filter(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, int lag_bytes, int grid,
                              cudaStream_t stream);
cudaError_t mmg_launch_resolve(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                               uint32_t *out_val, uint64_t capacity, cudaStream_t stream);
// one block of more than 128 segments: its phase prefix runs in two levels and needs MmgScratch::rangemap (+1 launch)
bool mmg_resolve_two_level(const MmgGeom &G);
// slice of a longer chain (mmg_chain_*): maps of the slice first; entry phases + resolve once G.entry is known
cudaError_t mmg_launch_chain_maps(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, cudaStream_t stream);
cudaError_t mmg_launch_chain_resolve(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                                     uint32_t *out_val, uint64_t capacity, cudaStream_t stream);
// sparse scans: the filter kernels resolve the engine blocks themselves when MmgScratch::fuse is set (no resolve launch);
// host_status[4] != 0 afterwards: a block was too dense, run mmg_launch_resolve over the same event lists
bool mmg_sparse_resolve_supported(const MmgGeom &G);
cudaError_t mmg_launch_scan(const uint32_t *counts, uint32_t n, uint64_t *bsum, uint64_t *bases, uint64_t *total,
                            cudaStream_t stream);
cudaError_t mmg_launch_emit(const MmgProgram &P, const MmgGeom &G, const MmgScratch &X, uint64_t *out_off,
                            uint32_t *out_val, cudaStream_t stream);
cudaError_t mmg_launch_generic_walk(const MmgProgram &P, const MmgGeom &G, uint32_t *counts, const uint64_t *bases,
                                    uint64_t *out_off, uint32_t *out_val, cudaStream_t stream);
cudaError_t mmg_launch_generic_walk_long(const MmgLongProgram &P, const MmgGeom &G, uint32_t *counts, const uint64_t *bases,
                                         uint64_t *out_off, uint32_t *out_val, cudaStream_t stream);
cudaError_t mmg_launch_generic_merge(uint32_t nblocks, const uint32_t *counts, const uint64_t *bases,
                                     const uint64_t *in_off, const uint32_t *in_val, uint64_t *out_off,
                                     uint32_t *out_val, cudaStream_t stream);

// distinct tables of a match list (unique.cu)
cudaError_t mmg_launch_unique_direct(const uint32_t *val, uint64_t n, uint32_t keymask, bool pack8, uint64_t *first,
                                     cudaStream_t stream);
cudaError_t mmg_launch_unique_hash(const uint32_t *val, uint64_t n, uint64_t *slots, uint32_t capmask, unsigned int *overflow,
                                   cudaStream_t stream);
cudaError_t mmg_launch_unique_collect(const uint64_t *slots, uint64_t nslots, bool low32, uint64_t *out, uint64_t *count,
                                      uint64_t capacity, cudaStream_t stream);

cudaError_t mmg_launch_synth(uint64_t *out, uint64_t nwords, uint64_t seed, uint64_t first_word, uint32_t byte_mask,
                             cudaStream_t stream);

#endif
