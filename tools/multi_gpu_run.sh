#!/bin/bash
# N = 2, 4, 8 bench lines + the merge correctness check on one 8-GPU box: tools/multi_gpu_run.sh r1e
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --steps 8 --warmup 3 > $out/bench_${tag}_n$n.json 2> $out/bench_${tag}_n$n.err
  tail -c 300 $out/bench_${tag}_n$n.json | head -c 10 > /dev/null
  python -c "import json; d=json.load(open('$out/bench_${tag}_n$n.json')); print('n=$n', d['value'], d['ms_per_step'], d.get('e2e'))"
done
$TR --nproc-per-node 8 --master-port 29555 tools/multi_gpu_check.py 2>&1 | tail -4
$TR --nproc-per-node 8 --master-port 29556 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 2>/dev/null | cut -c1-200
