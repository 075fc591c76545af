#!/bin/bash
# 1 GPU: tensor-map rows kernel (boxes of short runs) — parity tests + A/B timings against the per-run kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_rows_transpose_gpu.py tests/test_random_views_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_t17.log
tail -8 gpurun_out/r02_t17.log
timeout 200 python tools/transpose_cases.py --cases rows_256B_runs,rows_512B_runs,rows_1024B_runs,rows_2048B_runs,rows_3d_272B_runs > gpurun_out/r02_rows_cases.jsonl 2> gpurun_out/r02_rows_cases.err
echo "rc=$?"; tail -3 gpurun_out/r02_rows_cases.err
cat gpurun_out/r02_rows_cases.jsonl
