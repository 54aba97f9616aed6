#!/bin/bash
# GPU call 5 (r1h): branch-free triangle test (-DCRB_TRI_BRANCHFREE=1), A/B on config 2
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base tb1 base tb1; } > $out/ab5_r1h.txt 2>&1; cat $out/ab5_r1h.txt
