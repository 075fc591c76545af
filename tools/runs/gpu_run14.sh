#!/bin/bash
# 1 GPU: tensor-map TMA transpose kernel vs the LSU tiled transpose (correctness + CUDA-event time), then the view/transposes GPU tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 240 python tools/transpose_cases.py > gpurun_out/r02_transpose_cases.jsonl 2> gpurun_out/r02_transpose_cases.err
echo "rc=$?"; tail -5 gpurun_out/r02_transpose_cases.err
cat gpurun_out/r02_transpose_cases.jsonl
timeout 200 python -m pytest tests/test_rows_transpose_gpu.py tests/test_random_views_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_t14.log
tail -6 gpurun_out/r02_t14.log
