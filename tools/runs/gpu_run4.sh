# GPU run 4 (1 GPU): full GPU tests, kernel-case A/Bs, bench both arms with the NUMA defaults, ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t4_full.log 2>&1
tail -6 gpurun_out/r02_t4_full.log
for v in "occ2:TSNAP_B200_LSU_OCC=2" "occ3:TSNAP_B200_LSU_OCC=3" "rows128:TSNAP_B200_ROWS_MIN_RUN=128" "rows64:TSNAP_B200_ROWS_MIN_RUN=64"; do
  tag=${v%%:*}; kv=${v#*:}
  env $kv timeout 300 python tools/kernel_cases.py > gpurun_out/r02_kernel_cases_$tag.jsonl 2> gpurun_out/r02_kernel_cases_$tag.err
  echo "== $tag"; python -c "
import json,sys
for l in open('gpurun_out/r02_kernel_cases_$tag.jsonl'):
    d=json.loads(l); print('  %-28s %8.4f ms  %7.1f GB/s  %.3f' % (d['case'], d['kernel_ms'], d['gbs'], d['frac_of_measured_peak']))"
done
timeout 500 python bench.py --steps 10 --warmup 3 --trace-dir gpurun_out/r02_trace > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
cut -c1-300 gpurun_out/r02_bench_n1_c.json
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_n1_c.json 2> gpurun_out/r02_ref_n1_c.err
cut -c1-200 gpurun_out/r02_ref_n1_c.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
grep -c tsnap gpurun_out/r02_launches.csv
