# GPU run (8 GPUs): multi-GPU functional check, then C3 / C2 / C5 on 8 ranks, both arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/r02_t8.log 2>&1; tail -3 gpurun_out/r02_t8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { # cfg impl steps warmup extra...
  cfg=$1; impl=$2; steps=$3; warm=$4; shift 4
  out=gpurun_out/r02_${cfg}_n8_${impl}
  timeout 600 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 --config $cfg --impl $impl --steps $steps --warmup $warm "$@" > $out.json 2> $out.err
  echo "== $cfg $impl rc=$?"; cut -c1-230 $out.json
}
run c3 ours 4 2 --trace-dir gpurun_out/r02_trace --trace-tag c3
run c3 reference 3 1
run c2 ours 5 2
run c2 reference 3 1
run c5 ours 3 1
run c5 reference 2 1
