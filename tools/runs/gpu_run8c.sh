# GPU run (8 GPUs, short): multi-GPU functional check (read-once, error propagation, partition digest, folded key gather) + C2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 350 -p no:cacheprovider > gpurun_out/r02_t8c.log 2>&1; tail -3 gpurun_out/r02_t8c.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 8 --config c2 --steps 5 --warmup 2 > gpurun_out/r02_c2_n8_ours.json 2> gpurun_out/r02_c2_n8_ours.err
cut -c1-200 gpurun_out/r02_c2_n8_ours.json
