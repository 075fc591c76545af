# GPU run 10 (1 GPU): final state — smoke, full GPU tests, kernel cases, bench both arms (short reference)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t10_full.log 2>&1; tail -3 gpurun_out/r02_t10_full.log
timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_final.jsonl 2> gpurun_out/r02_kernel_cases_final.err
python -c "
import json
for l in open('gpurun_out/r02_kernel_cases_final.jsonl'):
    d=json.loads(l); print('  %-28s %8.4f ms  %7.1f GB/s  %.3f' % (d['case'], d['kernel_ms'], d['gbs'], d['frac_of_measured_peak']))"
timeout 500 python bench.py --steps 10 --warmup 3 --trace-dir gpurun_out/r02_trace > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_g.json')); print(d['value'], d['steps_ms'], d['take_blocking_ms']['async_take_returns_ms_each'], d['restore']['value'], d['drain']['link_starved_ms'], d['e2e_roofline']['frac'])"
