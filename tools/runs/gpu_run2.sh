# GPU run 2 (2 GPUs): full GPU test-suite with complete log, then the 2-GPU bench lines of every config, both arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x --deselect tests/test_native_gpu.py::test_column_shard_throughput_smoke > gpurun_out/r02_t2_full.log 2>&1
tail -5 gpurun_out/r02_t2_full.log
timeout 300 python -m pytest tests/test_native_gpu.py -m gpu -q -s -k column_shard -p no:cacheprovider > gpurun_out/r02_t2_colshard.log 2>&1; grep "column shard" gpurun_out/r02_t2_colshard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in c3 c2 c5 c4; do
  timeout 500 $TR --master-port 29601 bench.py --gpus 2 --config $cfg --steps 3 --warmup 2 > gpurun_out/r02_${cfg}_n2.json 2> gpurun_out/r02_${cfg}_n2.err
  timeout 500 $TR --master-port 29602 bench.py --impl reference --gpus 2 --config $cfg --steps 3 --warmup 1 > gpurun_out/r02_${cfg}_n2_ref.json 2> gpurun_out/r02_${cfg}_n2_ref.err
  echo "== $cfg"; cut -c1-260 gpurun_out/r02_${cfg}_n2.json; cut -c1-260 gpurun_out/r02_${cfg}_n2_ref.json
done
