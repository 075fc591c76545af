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
cudaError_t launch_lsu(const Member* d_members, const Tile* d_tiles, uint32_t ntiles, int sm_count,
                       cudaStream_t stream);
}  // namespace tsnap
