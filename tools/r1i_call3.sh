#!/bin/bash
# GPU call 3 (r1i): DMaterial alignment (dm4 = without) and the k_shade tile pipeline (sp1), A/B on config 2
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base dm4 sp1 base dm4 sp1; } > $out/ab_r1i3.txt 2>&1; cat $out/ab_r1i3.txt
