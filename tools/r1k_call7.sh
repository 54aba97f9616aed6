#!/bin/bash
# GPU call 7 (r1k): smoke + one 3-step bench after replacing the per-renderer cudaMemGetInfo by a tracked budget
out=gpurun_out; mkdir -p $out
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CRB_SUBMIT_DEBUG=1 CRB_BENCH_DEBUG=1 timeout 30 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/r1k_sub2.json 2> $out/r1k_sub2.err
{ awk '/crb submit/ { if ($4 + 0 > 1.5) print }' $out/r1k_sub2.err | tail -6; grep "e2e step ms" $out/r1k_sub2.err; python -c "import json; d=json.loads(open('$out/r1k_sub2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['steady_value'])"; } | tee $out/r1k_submit_debug2.txt
