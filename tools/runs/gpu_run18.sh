#!/bin/bash
# 1 GPU, end of round: full GPU suite, bench, smoke, kernel table refresh, ncu launch list of the tensor-map kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 260 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_t18_full.log
tail -3 gpurun_out/r02_t18_full.log
timeout 120 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
cut -c1-220 gpurun_out/r02_bench_n1_final.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r02_smoke_final.log 2>&1; tail -1 gpurun_out/r02_smoke_final.log
timeout 90 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2>/dev/null; wc -l gpurun_out/r02_kernel_cases_final.jsonl
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:tma --csv --log-file gpurun_out/r02_ncu_tma_kernels.csv \
  python tools/transpose_cases.py --cases fp32_8192x8192,fp32_16384x16384,bf16_16384x8192,fp64_4096x6144,rows_512B_runs --modes 1 --reps 1 > gpurun_out/r02_ncu_tma.log 2>&1
tail -2 gpurun_out/r02_ncu_tma.log
