#!/bin/bash
# GPU call 1 (r1k): bench line on the committed default (k_shade in CTAs of 128 threads)
out=gpurun_out; mkdir -p $out
timeout 170 python bench.py --steps 3 --warmup 3 > $out/r1k_bench.json 2> $out/r1k_bench.err; tail -2 $out/r1k_bench.err; cat $out/r1k_bench.json
