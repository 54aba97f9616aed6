#!/bin/bash
# GPU call 1 of the r1j refresh: parity tests (incl. compute-sanitizer on a small case), full profile refresh
out=gpurun_out; mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/pytest_r1j.log 2>&1; tail -5 $out/pytest_r1j.log
tools/refresh_profiles.sh r1j
{ python tools/bench_configs.py c2 --extended; python tools/bench_configs.py c5 --extended; python tools/bench_configs.py c1 --extended; } > $out/configs_ext_r1j.jsonl 2>> $out/bench_r1j.err
python -c "
import json
d=json.load(open('$out/bench_r1j.json')); print('value %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','steady_value','setup_ms')})
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
{ tools/ab_env.sh "CRB_TRACE_STEPS=2" "CRB_TRACE_STEPS=8" "CRB_COST_PRIM=0.6" "CRB_COST_PRIM=1.0"; } > $out/ab_r1j.txt 2>&1; cat $out/ab_r1j.txt
