# GPU run 3 (1 GPU): complete GPU test-suite, NUMA/ring sweep, kernel cases (CUDA events), ncu launch list + full captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t3_full.log 2>&1
tail -12 gpurun_out/r02_t3_full.log
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases.jsonl 2> gpurun_out/r02_kernel_cases.err; cat gpurun_out/r02_kernel_cases.jsonl | cut -c1-250
timeout 600 python tools/sweep_sink.py --reps 3 --set numa > gpurun_out/r02_sweep_numa.log 2> gpurun_out/r02_sweep_numa.err
timeout 300 python bench.py --steps 5 --warmup 2 --skip-cpu-baseline > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tsnap -o gpurun_out/r02_kernels python tools/kernel_cases.py --reps 1 > gpurun_out/r02_ncu_kernels.log 2>&1
tail -3 gpurun_out/r02_ncu_kernels.log
