#!/bin/bash
# GPU call 7 (r1h): k_shade / k_shade_ext with all slot-indexed loads issued together (-DCRB_SHADE_GROUPED_LOADS=1), A/B
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base gl1 base gl1
  for v in base gl1; do CRENDER_B200_LIB=crender_b200/_variants/libv_$v.so python tools/bench_configs.py c2 --extended | cut -c1-20,180-; done; } > $out/ab7_r1h.txt 2>&1; cat $out/ab7_r1h.txt
