#!/bin/bash
# GPU call 2 (r1k): where the e2e loop spends its time at 3 and at 16 steps (per-step wall marks)
out=gpurun_out; mkdir -p $out
for k in 3 3 16; do
  CRB_BENCH_DEBUG=1 timeout 120 python bench.py --steps $k --warmup 3 --no-cpu-baseline --no-roofline > $out/r1k_e2e_$k.json 2> $out/r1k_e2e_$k.err
  grep "e2e" $out/r1k_e2e_$k.err | tail -5
  python -c "import json,sys; d=json.loads(open('$out/r1k_e2e_$k.json').read().strip().splitlines()[-1]); print($k, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['steady_value'], d['e2e']['setup_ms'])"
done
