#!/bin/bash
# GPU call 2 (r1i): early pop as predicated in-place local loads (-DCRB_EARLY_POP=2) vs the C++ form (=1): parity + A/B
out=gpurun_out; mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/pytest_r1i2.log 2>&1; tail -3 $out/pytest_r1i2.log
{ tools/ab.sh ep1 ep2 ep1 ep2; } > $out/ab_r1i2.txt 2>&1; cat $out/ab_r1i2.txt
python tools/setup_breakdown.py 2>&1 | tail -3
