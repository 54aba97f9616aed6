#!/bin/bash
# e2e per-step timing check (CRB_BENCH_DEBUG prints the wall time of every e2e step)
out=gpurun_out; mkdir -p $out
for i in 1 2 3 4; do CRB_BENCH_DEBUG=1 python bench.py --no-cpu-baseline --no-roofline 2> $out/e2e_dbg_$i.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.2f'%(d['value'], d['ms_per_step']), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','steady_value','setup_ms','upload_ms')})"; grep "e2e step" $out/e2e_dbg_$i.err; done
