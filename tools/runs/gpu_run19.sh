#!/bin/bash
# 1 GPU, last seconds of the budget: the TMA transpose tests after the load-stage release moved behind the stores
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_rows_transpose_gpu.py -q -m gpu -p no:cacheprovider -k "tma" 2>&1 | tail -4 > gpurun_out/r02_t19.log
tail -3 gpurun_out/r02_t19.log
timeout 30 python tools/transpose_cases.py --cases fp32_8192x8192,bf16_16384x8192 --modes 1 --reps 3 2>/dev/null | tee gpurun_out/r02_transpose_cases_final.jsonl
