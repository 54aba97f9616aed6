#!/bin/bash
# GPU call 6 (r1k): where the slow first submit goes (CRB_SUBMIT_DEBUG), default vs CRB_LMEM_MAX=1
out=gpurun_out; mkdir -p $out
for v in 0 1 0 1 0 1; do
  echo "== CRB_LMEM_MAX=$v"
  CRB_LMEM_MAX=$v CRB_SUBMIT_DEBUG=1 CRB_BENCH_DEBUG=1 timeout 40 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/r1k_sub.json 2> $out/r1k_sub.err
  awk '/crb submit/ { split($4, a, " "); if ($4 + 0 > 1.5) print }' $out/r1k_sub.err | tail -6
  grep "e2e step ms" $out/r1k_sub.err
done 2>&1 | tee $out/r1k_submit_debug.txt
