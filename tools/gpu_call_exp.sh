#!/bin/bash
mkdir -p gpurun_out
TAG=$1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step e2e',d['e2e']['value'],d['clocks'], d.get('variants'), r.get('traffic'), d.get('cpu_baseline'))
"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 | tail -c 700
