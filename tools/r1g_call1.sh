#!/bin/bash
# GPU call 1 of the r1g refresh: parity tests, headline bench, node-test variants A/B, extended-mode configs
out=gpurun_out; mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/pytest_r1g.log 2>&1; tail -5 $out/pytest_r1g.log
python bench.py > $out/bench_r1g.json 2> $out/bench_r1g.err; tail -c 400 $out/bench_r1g.json
{ tools/ab_env.sh "CRB_BASE=1"; tools/ab.sh prmt1 prmt2; tools/ab_env.sh "CRB_BASE=2"; } > $out/ab_r1g.txt 2>&1; cat $out/ab_r1g.txt
{ python tools/bench_configs.py c2; python tools/bench_configs.py c2 --extended; python tools/bench_configs.py c2 --extended --sort;
  python tools/bench_configs.py c5 --extended; python tools/bench_configs.py c1 --extended; } > $out/configs_ext_r1g.jsonl 2>> $out/bench_r1g.err
cat $out/configs_ext_r1g.jsonl
