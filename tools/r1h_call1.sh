#!/bin/bash
# GPU call 1 of the r1h refresh: parity tests, full profile refresh (bench, reference arm, ncu), extended-mode
# configs, e2e with the blocking read for comparison, refill-granularity re-check
out=gpurun_out; mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/pytest_r1h.log 2>&1; tail -5 $out/pytest_r1h.log
tools/refresh_profiles.sh r1h
python bench.py --steps 16 --no-cpu-baseline --no-roofline --e2e-blocking-read > $out/bench_blockingread_r1h.json 2>> $out/bench_r1h.err
{ python tools/bench_configs.py c2 --extended; python tools/bench_configs.py c2 --extended --sort;
  python tools/bench_configs.py c5 --extended; python tools/bench_configs.py c1 --extended; } > $out/configs_ext_r1h.jsonl 2>> $out/bench_r1h.err
cat $out/configs_ext_r1h.jsonl
{ tools/ab.sh base occ3 base occ3; tools/ab_env.sh "CRB_TRACE_STEPS=2" "CRB_TRACE_STEPS=8" "CRB_COST_PRIM=0.6" "CRB_COST_PRIM=1.0" "CRB_COST_PRIM=1.4"; } > $out/ab_r1h.txt 2>&1; cat $out/ab_r1h.txt
python -c "
import json
for f in ('bench_r1h','bench_blockingread_r1h'):
    d=json.load(open('$out/'+f+'.json')); print(f, 'value %.0f'%d['value'], 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','steady_value','setup_ms','read_back')})
"
