#!/bin/bash
# GPU call 2 of the r1h refresh (2 GPUs): bench at N=2 with the overlapped and the blocking merged-buffer read-back,
# merge correctness check
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
$TR --master-port 29502 bench.py --gpus 2 --steps 8 --warmup 3 > $out/bench_r1h_n2.json 2> $out/bench_r1h_n2.err
$TR --master-port 29503 bench.py --gpus 2 --steps 8 --warmup 3 --e2e-blocking-read > $out/bench_r1h_n2_blocking.json 2>> $out/bench_r1h_n2.err
for f in bench_r1h_n2 bench_r1h_n2_blocking; do python -c "import json; d=json.load(open('$out/$f.json')); print('$f', d['value'], d['ms_per_step'], d.get('e2e'))"; done
$TR --master-port 29555 tools/multi_gpu_check.py 2>&1 | tail -4
tail -5 $out/bench_r1h_n2.err
