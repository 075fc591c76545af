# GPU run 11 (1 GPU): tests after the host-budget / slot accounting change + short bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r02_t11_full.log 2>&1; tail -4 gpurun_out/r02_t11_full.log
timeout 400 python bench.py --steps 6 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_bench_n1_h.json 2> gpurun_out/r02_bench_n1_h.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_h.json')); print(d['value'], d['steps_ms'], d['take_blocking_ms']['async_take_returns_ms_each'], d['restore']['value'], d['engine_step'].get('max_slots_in_flight'))"
