#!/bin/bash
# 1 GPU: transpose tile geometry x enumeration-order A/B, plus the transpose/rows GPU tests on the new tile decode
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/transpose_tune.py > gpurun_out/r02_transpose_tune.jsonl 2> gpurun_out/r02_transpose_tune.err
tail -3 gpurun_out/r02_transpose_tune.err
timeout 150 python -m pytest tests/test_rows_transpose_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/r02_t12.log
tail -3 gpurun_out/r02_t12.log
wc -l gpurun_out/r02_transpose_tune.jsonl
