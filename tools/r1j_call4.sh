#!/bin/bash
# GPU call 3 of the r1j refresh (8 GPUs): bench at N=4 and N=8, merge correctness check at 8
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --steps 8 --warmup 3 > $out/bench_r1j_n$n.json 2> $out/bench_r1j_n$n.err
  python -c "import json; d=json.load(open('$out/bench_r1j_n$n.json')); print('n=$n', d['value'], d['ms_per_step'], d.get('e2e'))"
done
$TR --nproc-per-node 8 --master-port 29555 tools/multi_gpu_check.py 2>&1 | grep -E "relRMSE|MULTI_GPU"
