# GPU run 8 (1 GPU): tests after the transpose-build change, kernel cases, C1 on the GPU box's host, short bench (link_starved_ms)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t8_full.log 2>&1; tail -4 gpurun_out/r02_t8_full.log
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2> gpurun_out/r02_kernel_cases_final.err
python -c "
import json
for l in open('gpurun_out/r02_kernel_cases_final.jsonl'):
    d=json.loads(l); print('  %-28s %8.4f ms  %7.1f GB/s  %.3f' % (d['case'], d['kernel_ms'], d['gbs'], d['frac_of_measured_peak']))"
for impl in ours reference; do timeout 300 python bench.py --config c1 --impl $impl --steps 5 --warmup 2 > gpurun_out/r02_c1_gpubox_$impl.json 2> gpurun_out/r02_c1_gpubox_$impl.err; cut -c1-160 gpurun_out/r02_c1_gpubox_$impl.json; done
timeout 400 python bench.py --steps 6 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_f.json')); print(d['value'], d['steps_ms'], d['drain'], d['restore']['value'])"
timeout 600 ncu --set full --clock-control none -k regex:tsnap_lsu -o gpurun_out/r02_kernels_lsu python tools/kernel_cases.py --reps 0 > gpurun_out/r02_ncu_kernels_lsu.log 2>&1
python tools/ncu_summarize.py gpurun_out/r02_kernels_lsu.ncu-rep gpurun_out/r02_ncu_lsu odd_align transpose_fp32 transpose_bf16 cast strided_128B > /dev/null 2>&1; cat gpurun_out/r02_ncu_lsu_table.md
rm -f gpurun_out/r02_kernels_lsu.ncu-rep
