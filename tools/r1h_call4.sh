#!/bin/bash
# GPU call 4 (r1h): register-cached stack top (-DCRB_STACK_CACHE=1) and small look-ahead work reservation, A/B on config 2
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base sc1 base sc1; tools/ab_env.sh "CRB_TRACE_CHUNK=4" "CRB_TRACE_CHUNK=8" "CRB_TRACE_CHUNK=16"; } > $out/ab4_r1h.txt 2>&1; cat $out/ab4_r1h.txt
