# GPU run 7 (1 GPU): tests after the transpose/tile-prefetch change, kernel cases, short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t7_full.log 2>&1; tail -4 gpurun_out/r02_t7_full.log
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2> gpurun_out/r02_kernel_cases_final.err
python -c "
import json
for l in open('gpurun_out/r02_kernel_cases_final.jsonl'):
    d=json.loads(l); print('  %-28s %8.4f ms  %7.1f GB/s  %.3f' % (d['case'], d['kernel_ms'], d['gbs'], d['frac_of_measured_peak']))"
timeout 400 python bench.py --steps 6 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_e.json')); print(d['value'], d['steps_ms'], d['take_blocking_ms']['async_take_returns_ms_each'], d['restore']['value'])"
timeout 600 ncu --set full --clock-control none -k regex:tsnap_lsu -o gpurun_out/r02_kernels_lsu python tools/kernel_cases.py --reps 0 > gpurun_out/r02_ncu_kernels_lsu.log 2>&1
python tools/ncu_summarize.py gpurun_out/r02_kernels_lsu.ncu-rep gpurun_out/r02_ncu_lsu odd_align transpose_fp32 transpose_bf16 cast strided_128B > /dev/null 2>&1; cat gpurun_out/r02_ncu_lsu_table.md
rm -f gpurun_out/r02_kernels_lsu.ncu-rep
