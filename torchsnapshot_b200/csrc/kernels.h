#pragma once
#include <cuda_runtime.h>

#include "plan.h"

namespace tsnap {
// one-time attribute setup (opt-in dynamic shared memory for the bulk kernel)
cudaError_t init_kernels();
cudaError_t launch_bulk(const Member* d_members, const Tile* d_tiles, uint32_t ntiles, int sm_count,
                        cudaStream_t stream);
cudaError_t launch_rows(const Member* d_members, const Tile* d_tiles, uint32_t ntiles, int sm_count,
                        cudaStream_t stream);
// Which build of the LSU kernel a tile list runs on (register bound = resident CTAs per SM):
//   kLsuDefault   3 CTAs/SM, 80 registers: unaligned dense runs 0.85, casts 0.93 of the measured HBM peak (0.74 / 0.85 on the 2-CTA build)
//   kLsuStrided   2 CTAs/SM, 128 registers: the strided gather keeps 8 independent 16 B loads per thread in flight without
//                 spilling (0.66 vs 0.47 on 128 B runs)
//   kLsuTranspose 6 CTAs/SM, 40 registers: 16 KiB tiles, more CTAs in different phases of the load -> shared -> store cycle
//                 (the fallback for transposes the tensor-map kernel cannot take)
enum LsuVariant { kLsuDefault = 0, kLsuStrided = 1, kLsuTranspose = 2 };
cudaError_t launch_lsu(const Member* d_members, const Tile* d_tiles, uint32_t ntiles, int sm_count,
                       cudaStream_t stream, int variant);

// kModeTransposeTma (transpose_tma.cu): tensor-map TMA tile loads/stores around a register block transposition.
// make_tma_pair encodes the two tensor maps of a kModeTranspose member (false: the member stays on the LSU transpose —
// 1-byte elements, bases or strides that are not 16 B multiples, more than 3 extra dims, TSNAP_B200_TMA_TRANSPOSE=0,
// or a driver without cuTensorMapEncodeTiled).
bool transpose_tma_enabled();
bool make_tma_pair(const Member& m, TmaPair* out, uint32_t* variant);  // variant = tile shape, for Member.shift bits 16-17
// kModeRowsTma: the same for kModeRows members with runs <= 1 KiB (TSNAP_B200_TMA_ROWS=0 keeps them on the per-run kernel)
bool make_rows_tma_pair(const Member& m, TmaPair* out);
cudaError_t launch_rows_tma(const Member* d_members, const Tile* d_tiles, const TmaPair* d_maps, uint32_t ntiles, int sm_count,
                            cudaStream_t stream);
cudaError_t init_transpose_tma();  // both tensor-map kernels
cudaError_t launch_transpose_tma(const Member* d_members, const Tile* d_tiles, const TmaPair* d_maps, uint32_t ntiles, int sm_count,
                                 cudaStream_t stream);
}  // namespace tsnap
