#!/bin/bash
# like ab.sh but "variant:ENV=val,ENV2=val" entries set environment knobs for that run
for e in "$@"; do
  v=${e%%:*}; envs=""; [[ "$e" == *:* ]] && envs=$(echo "${e#*:}" | tr ',' ' ')
  env $envs CRENDER_B200_LIB=crender_b200/_variants/libv_$v.so python bench.py --steps 6 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$e', 'value %.0f'%d['value'], 'trace %.0f shadow %.0f'%(r['mrays_s_trace_kernel'], r['mrays_s_shadow_kernel']), {k:round(v,1) for k,v in r['kernel_ms'].items()})"
done
