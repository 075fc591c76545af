mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_env.txt 2>&1
uname -r >> gpurun_out/r02_env.txt; df -h /tmp /dev/shm >> gpurun_out/r02_env.txt; grep -E " / | /tmp " /proc/mounts >> gpurun_out/r02_env.txt
cat /sys/kernel/mm/transparent_hugepage/shmem_enabled >> gpurun_out/r02_env.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_t1.log
timeout 600 python tools/sweep_sink.py --reps 3 > gpurun_out/r02_sweep.log 2> gpurun_out/r02_sweep.err
timeout 400 python bench.py --steps 5 --warmup 2 --trace-dir gpurun_out/r02_trace > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_n1.json 2> gpurun_out/r02_ref_n1.err
tail -5 gpurun_out/r02_t1.log; tail -3 gpurun_out/r02_sweep.log | cut -c1-300; cut -c1-400 gpurun_out/r02_bench_n1.json; cut -c1-300 gpurun_out/r02_ref_n1.json
