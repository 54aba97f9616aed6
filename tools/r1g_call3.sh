#!/bin/bash
out=gpurun_out; mkdir -p $out
{ tools/ab.sh p11 fused1 l1max fused1l1 fused1p0 prmt12 p11 fused1; } > $out/ab3_r1g.txt 2>&1; cat $out/ab3_r1g.txt
