#!/bin/bash
# 1 GPU: TMA transpose kernel — tile-shape variants, new parity tests, ncu --set full capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_rows_transpose_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_t15.log
tail -8 gpurun_out/r02_t15.log
timeout 200 python tools/transpose_cases.py > gpurun_out/r02_transpose_cases.jsonl 2> gpurun_out/r02_transpose_cases.err
echo "rc=$?"; tail -3 gpurun_out/r02_transpose_cases.err
cat gpurun_out/r02_transpose_cases.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tsnap_transpose_tma -o gpurun_out/r02_tt \
  python tools/transpose_cases.py --cases fp32_8192x8192,bf16_16384x8192 --modes 1 --reps 1 > gpurun_out/r02_ncu_tt.log 2>&1
ncu -i gpurun_out/r02_tt.ncu-rep --page raw --csv > gpurun_out/r02_ncu_tt_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_tt.ncu-rep --page details --csv > gpurun_out/r02_ncu_tt_details.csv 2>/dev/null
ls -la gpurun_out/r02_tt.ncu-rep
if [ $(stat -c %s gpurun_out/r02_tt.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/r02_tt.ncu-rep; fi
tail -3 gpurun_out/r02_ncu_tt.log
