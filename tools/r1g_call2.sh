#!/bin/bash
out=gpurun_out; mkdir -p $out
{ tools/ab.sh prmt11 prmt2 prmt12 prmt1; } > $out/ab2_r1g.txt 2>&1; cat $out/ab2_r1g.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline"
CRENDER_B200_LIB=crender_b200/_variants/libv_prmt2.so ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 9 --launch-count 2 -f -o $out/prof_trace_prmt2 $B > $out/ncu_t.log 2>&1
tail -3 $out/ncu_t.log
