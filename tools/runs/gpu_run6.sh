# GPU run 6 (1 GPU): final N=1 artefacts — bench both arms, trace diagnosis, kernel cases, launch list, ncu full capture
# (exported to CSV on the box; the .ncu-rep is dropped if it would blow the 64 MiB return limit)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t6_full.log 2>&1; tail -4 gpurun_out/r02_t6_full.log
timeout 300 python tools/sweep_sink.py --set trace > gpurun_out/r02_trace_diag.jsonl 2> gpurun_out/r02_trace_diag.err
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2> gpurun_out/r02_kernel_cases_final.err
timeout 500 python bench.py --steps 10 --warmup 3 --trace-dir gpurun_out/r02_trace > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; cut -c1-200 gpurun_out/r02_bench_n1_d.json
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_n1_d.json 2> gpurun_out/r02_ref_n1_d.err; cut -c1-200 gpurun_out/r02_ref_n1_d.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; grep -c tsnap gpurun_out/r02_launches.csv
timeout 600 ncu --set full --clock-control none -k regex:tsnap -o gpurun_out/r02_kernels_final python tools/kernel_cases.py --reps 0 > gpurun_out/r02_ncu_kernels_final.log 2>&1; tail -2 gpurun_out/r02_ncu_kernels_final.log
python tools/ncu_summarize.py gpurun_out/r02_kernels_final.ncu-rep gpurun_out/r02_ncu_kernels dense_bulk column_512B column_8KiB odd_align transpose_fp32 transpose_bf16 cast strided_128B > /dev/null 2>&1
ncu -i gpurun_out/r02_kernels_final.ncu-rep --page details --csv > gpurun_out/r02_ncu_kernels_details.csv 2>/dev/null
ls -la gpurun_out/ | head -40
sz=$(stat -c %s gpurun_out/r02_kernels_final.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/r02_kernels_final.ncu-rep; echo "dropped ncu-rep ($sz bytes), CSV exports kept"; fi
du -sh gpurun_out
