#!/bin/bash
# final r1j bench line (default arguments) + reference arm
out=gpurun_out; mkdir -p $out
python bench.py > $out/bench_r1j.json 2> $out/bench_r1j.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_r1j.json 2>> $out/bench_r1j.err
python -c "
import json
d=json.load(open('$out/bench_r1j.json')); print('value %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','steady_value','setup_ms')}, 'launches', d['gpu_launches'], d['clocks'])
"
