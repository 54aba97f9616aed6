#!/bin/bash
# GPU call 5 (r1k): e2e at 3 steps with the interpreter's collector kept out of the timed region, four runs
out=gpurun_out; mkdir -p $out
for i in 1 2 3 4; do
  CRB_BENCH_DEBUG=1 timeout 60 python bench.py --steps 3 --warmup 3 --cpu-seconds 2 > $out/r1k_gc_$i.json 2> $out/r1k_gc_$i.err
  grep "e2e step ms" $out/r1k_gc_$i.err
  python -c "import json,sys; d=json.loads(open('$out/r1k_gc_$i.json').read().strip().splitlines()[-1]); print($i, d['value'], d['e2e']['value'], d['e2e']['steady_value'], d['e2e']['setup_ms'])"
done
