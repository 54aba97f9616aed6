#!/bin/bash
# One-call refresh of everything under profiles/ for a round tag: tools/refresh_profiles.sh r1e
# (run on the GPU box through gpurun; raw ncu files land in gpurun_out/, summaries are made afterwards with
#  tools/ncu_summary.py in the development container)
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 8 --launch-count 8 -f -o $out/prof_trace_$tag $B > $out/ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_shadow|k_shade' --launch-skip 16 --launch-count 6 -f -o $out/prof_shade_$tag $B > $out/ncu_s.log 2>&1
for c in mem build c1 c4 c5; do python tools/bench_configs.py $c; done > $out/configs_$tag.jsonl 2>> $out/bench_$tag.err
python tools/bench_configs.py rays >> $out/configs_$tag.jsonl 2>> $out/bench_$tag.err
python tools/bench_configs.py rays --no-ground >> $out/configs_$tag.jsonl 2>> $out/bench_$tag.err
tail -c 600 $out/bench_$tag.json; tail -3 $out/ncu_t.log
