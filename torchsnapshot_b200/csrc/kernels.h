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
// strided = true: the build of the LSU kernel bounded for 2 CTAs per SM (128 registers: the strided gather keeps 8
// independent 16 B loads per thread in flight without spilling; 0.65 vs 0.47 of peak on 128 B runs).  All other modes
// run the 3-CTA build (80 registers: 0.84 / 0.92 of peak on unaligned dense runs / casts vs 0.74 / 0.85).
cudaError_t launch_lsu(const Member* d_members, const Tile* d_tiles, uint32_t ntiles, int sm_count,
                       cudaStream_t stream, bool strided);
}  // namespace tsnap
