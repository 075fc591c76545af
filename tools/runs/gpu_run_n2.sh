# GPU run (2 GPUs): multi-GPU pytest (partitioner digest path, read-once restore under NCCL), C3 and C2 at N=2 both arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/r02_t_n2.log 2>&1; tail -2 gpurun_out/r02_t_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { cfg=$1; impl=$2; steps=$3; warm=$4; shift 4
  out=gpurun_out/r02_${cfg}_n2_${impl}
  timeout 600 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --config $cfg --impl $impl --steps $steps --warmup $warm "$@" > $out.json 2> $out.err
  echo "== $cfg $impl rc=$?"; cut -c1-230 $out.json; }
run c3 ours 5 2
run c3 reference 3 1
run c2 ours 5 2
