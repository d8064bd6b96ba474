#!/bin/bash
# 2-GPU validation: distributed sampling check + the bench exactly as the driver launches it (no cpu leg to save box time)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/dist_sample_check.py 2>&1 | grep "rank\|Error\|error" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_2gpu_r02b.json 2> gpurun_out/bench_2gpu_r02b.err
tail -c 1500 gpurun_out/bench_2gpu_r02b.json; tail -3 gpurun_out/bench_2gpu_r02b.err
