#!/bin/bash
# A/B of environment knobs on the GPU box with the in-tree library: tools/ab_env.sh "CRB_X=0" "CRB_X=1" ...
for v in "$@"; do
  env $v python bench.py --steps 6 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', 'value %.0f'%d['value'], 'trace %.0f shadow %.0f'%(r['mrays_s_trace_kernel'], r['mrays_s_shadow_kernel']), {k:round(v,1) for k,v in r['kernel_ms'].items()})"
done
