#!/bin/bash
# GPU call 6 (r1j): k_shade CTA size 128 (default now) vs 64; parity tests on the new default
out=gpurun_out; mkdir -p $out
{ tools/ab.sh base sb64 base sb64; } > $out/ab_r1j6.txt 2>&1; cat $out/ab_r1j6.txt
( time python -m pytest tests -m gpu -x -q ) > $out/pytest_r1j6.log 2>&1; tail -3 $out/pytest_r1j6.log
