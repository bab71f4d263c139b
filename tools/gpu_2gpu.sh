#!/bin/bash
# two-GPU check of the driver's launch lines (both arms) + the world-size-2 NCCL path
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"
tail -c 400 gpurun_out/bench_2gpu.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','n_gpus','ms_per_step','scaling','e2e','clocks','parity']})"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "reference arm 2gpu exit $?"
tail -c 300 gpurun_out/bench_2gpu_ref.err; tail -1 gpurun_out/bench_2gpu_ref.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --scaling strong > gpurun_out/bench_2gpu_strong.json 2> gpurun_out/bench_2gpu_strong.err; echo "strong exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu_strong.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ['value','n_gpus','ms_per_step','scaling']}, d['config'].get('batch_per_gpu'))"
