#!/bin/bash
# N=2: multi-GPU functional check only (read-once accounting + missing-file negative test with chunked entries)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_multigpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_t2d.log
tail -5 gpurun_out/r02_t2d.log
