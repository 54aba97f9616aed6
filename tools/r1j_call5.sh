#!/bin/bash
# GPU call 5 (r1j): CTA size of k_shade (its queue push is CTA-collective), A/B on config 2
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base sb128 sb512 base sb128 sb512; } > $out/ab_r1j5.txt 2>&1; cat $out/ab_r1j5.txt
