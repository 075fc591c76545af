# GPU run 9 (1 GPU): tests + kernel cases after the conflict-free transpose paths; ncu of the LSU builds
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t9_full.log 2>&1; tail -4 gpurun_out/r02_t9_full.log
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2> gpurun_out/r02_kernel_cases_final.err
python -c "
import json
for l in open('gpurun_out/r02_kernel_cases_final.jsonl'):
    d=json.loads(l); print('  %-28s %8.4f ms  %7.1f GB/s  %.3f' % (d['case'], d['kernel_ms'], d['gbs'], d['frac_of_measured_peak']))"
timeout 600 ncu --set full --clock-control none -k regex:tsnap_lsu -o gpurun_out/r02_kernels_lsu python tools/kernel_cases.py --reps 0 > gpurun_out/r02_ncu_kernels_lsu.log 2>&1
python tools/ncu_summarize.py gpurun_out/r02_kernels_lsu.ncu-rep gpurun_out/r02_ncu_lsu odd_align transpose_fp32 transpose_bf16 cast strided_128B > /dev/null 2>&1; cat gpurun_out/r02_ncu_lsu_table.md
ncu -i gpurun_out/r02_kernels_lsu.ncu-rep --page details --csv > gpurun_out/r02_ncu_lsu_details.csv 2>/dev/null
rm -f gpurun_out/r02_kernels_lsu.ncu-rep
