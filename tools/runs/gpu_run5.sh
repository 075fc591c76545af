# GPU run 5 (1 GPU): where do slow takes come from (traced), full tests, kernel cases + ncu full capture, final bench + launch list
mkdir -p gpurun_out
timeout 300 python tools/sweep_sink.py --set trace > gpurun_out/r02_trace_diag.jsonl 2> gpurun_out/r02_trace_diag.err; cut -c1-420 gpurun_out/r02_trace_diag.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t5_full.log 2>&1; tail -4 gpurun_out/r02_t5_full.log
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2> gpurun_out/r02_kernel_cases_final.err
python -c "
import json
for l in open('gpurun_out/r02_kernel_cases_final.jsonl'):
    d=json.loads(l); print('  %-28s %8.4f ms  %7.1f GB/s  %.3f' % (d['case'], d['kernel_ms'], d['gbs'], d['frac_of_measured_peak']))"
timeout 500 python bench.py --steps 10 --warmup 3 --trace-dir gpurun_out/r02_trace > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; cut -c1-200 gpurun_out/r02_bench_n1_d.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_d.json')); print(d['steps_ms'], d['take_blocking_ms']['async_take_returns_ms_each'], d['restore']['value'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tsnap -o gpurun_out/r02_kernels_final python tools/kernel_cases.py --reps 1 > gpurun_out/r02_ncu_kernels_final.log 2>&1; tail -2 gpurun_out/r02_ncu_kernels_final.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; grep -c tsnap gpurun_out/r02_launches.csv
