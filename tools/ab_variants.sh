#!/bin/bash
# A/B of library variants / tuning knobs on the GPU box: tools/ab_variants.sh <tag> "<name>|<lib or ->|<ENV=V ...>" ...
# one short headline bench per entry (device-timed value + per-kernel rates from the instrumented pass)
tag=$1; shift
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  IFS='|' read -r name lib envs <<< "$spec"
  [ "$lib" = "-" ] && libenv="" || libenv="CRENDER_B200_LIB=$PWD/crender_b200/_variants/libv_$lib.so"
  env $libenv $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $out/${tag}_ab_$name.json 2> $out/${tag}_ab_$name.err
  python - "$name" $out/${tag}_ab_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); r = d.get("roofline", {})
    print("%-14s value %8.1f Mrays/s  ms/step %6.2f  trace %7.1f shadow %7.1f  nodes/q %.3f sh_nodes/q %.3f  share %s" % (sys.argv[1], d["value"], d["ms_per_step"], r.get("mrays_s_trace_kernel") or 0, r.get("mrays_s_shadow_kernel") or 0, r.get("nodes_per_query") or 0, r.get("shadow_nodes_per_query") or 0, {k: round(v, 3) for k, v in (r.get("kernel_ms_share") or {}).items() if v}))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
