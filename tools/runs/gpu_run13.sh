#!/bin/bash
# 1 GPU: which side's run length bounds the tiled transpose (tall-skinny shapes: one side of each tile is one contiguous block)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/transpose_tune.py --geoms 0 --panels 1,4 --cases fp32_Nx64_src_dense,fp32_64xN_dst_dense,fp32_Nx128,fp32_128xN,fp32_Nx1024,fp32_1024xN,fp32_8192x8192 > gpurun_out/r02_transpose_sides.jsonl 2> gpurun_out/r02_transpose_sides.err
tail -3 gpurun_out/r02_transpose_sides.err
cat gpurun_out/r02_transpose_sides.jsonl
