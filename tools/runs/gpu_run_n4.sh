# GPU run (4 GPUs): multi-GPU pytest, C4 (async_take under DDP training) and C3, both arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/r02_t_n4.log 2>&1; tail -2 gpurun_out/r02_t_n4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { cfg=$1; impl=$2; steps=$3; warm=$4; shift 4
  out=gpurun_out/r02_${cfg}_n4_${impl}
  timeout 600 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 4 --config $cfg --impl $impl --steps $steps --warmup $warm "$@" > $out.json 2> $out.err
  echo "== $cfg $impl rc=$?"; cut -c1-230 $out.json; }
run c4 ours 4 2
run c4 reference 3 1
run c3 ours 4 2
run c3 reference 3 1
