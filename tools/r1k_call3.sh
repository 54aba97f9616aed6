#!/bin/bash
# GPU call 3 (r1k): full bench line (roofline + e2e + cpu baseline) at 3 and at the default 16 steps, with the e2e per-step marks
out=gpurun_out; mkdir -p $out
for k in 3 16; do
  CRB_BENCH_DEBUG=1 timeout 150 python bench.py --steps $k --warmup 3 > $out/r1k_full_$k.json 2> $out/r1k_full_$k.err
  grep "e2e step ms" $out/r1k_full_$k.err
  python -c "import json,sys; d=json.loads(open('$out/r1k_full_$k.json').read().strip().splitlines()[-1]); print($k, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['steady_value'], d['e2e']['setup_ms'], d['roofline']['frac'])"
done
