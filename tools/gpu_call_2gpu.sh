#!/bin/bash
# 2-GPU check of the driver's launch line: one rank per GPU, weights broadcast once, independent chain shards.
mkdir -p gpurun_out
TAG=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_2gpu.json 2> gpurun_out/bench_${TAG}_2gpu.err; echo "2gpu bench exit $?"
tail -c 1500 gpurun_out/bench_${TAG}_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-batch 4 > gpurun_out/bench_${TAG}_2gpu_ref.json 2> gpurun_out/bench_${TAG}_2gpu_ref.err; echo "2gpu ref exit $?"
tail -c 600 gpurun_out/bench_${TAG}_2gpu_ref.json
tail -5 gpurun_out/bench_${TAG}_2gpu.err
