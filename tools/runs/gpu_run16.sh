#!/bin/bash
# 1 GPU: TMA transpose kernel with a dedicated producer warp (mbarrier pipeline, no CTA-wide barriers): parity tests + timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_rows_transpose_gpu.py tests/test_random_views_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_t16.log
tail -8 gpurun_out/r02_t16.log
timeout 200 python tools/transpose_cases.py > gpurun_out/r02_transpose_cases.jsonl 2> gpurun_out/r02_transpose_cases.err
echo "rc=$?"; tail -3 gpurun_out/r02_transpose_cases.err
cat gpurun_out/r02_transpose_cases.jsonl
