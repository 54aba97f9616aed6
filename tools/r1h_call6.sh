#!/bin/bash
# GPU call 6 (r1h): pop before the leaf phase (-DCRB_EARLY_POP=1) on top of the branch-free triangle test, A/B on config 2
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base ep1 base ep1; } > $out/ab6_r1h.txt 2>&1; cat $out/ab6_r1h.txt
