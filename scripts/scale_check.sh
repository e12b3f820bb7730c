#!/bin/bash
# runs bench.py at N = 2, 4, 8 ranks back to back (as the driver's scaling step does); each with a hard timeout
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 1500 --warmup 100 2>&1 | tail -1 > gpurun_out/bench_r1_n$n.json
  python -c "import json; d=json.load(open('gpurun_out/bench_r1_n$n.json')); print($n, round(d['value']), round(d['ms_per_step']*1000,1), 'us/step', d['clocks'])" || tail -3 gpurun_out/bench_r1_n$n.json
done
