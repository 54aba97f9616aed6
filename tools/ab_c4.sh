#!/bin/bash
# A/B of library variants on config 4 (instanced terrain + city, two-level traversal): tools/ab_c4.sh <tag> "<name>|<lib or ->|<ENV=V ...>" ...
tag=$1; shift
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  IFS='|' read -r name lib envs <<< "$spec"
  [ "$lib" = "-" ] && libenv="" || libenv="CRENDER_B200_LIB=$PWD/crender_b200/_variants/libv_$lib.so"
  env $libenv $envs timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --only c4 --c4-spp 128 > $out/${tag}_c4_$name.json 2> $out/${tag}_c4_$name.err
  python - "$name" $out/${tag}_c4_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); s = d["strong_c4"]; r = s.get("roofline", {}); f = s.get("flattened", {})
    print("%-12s c2 %7.1f | c4 two-level %7.1f Mrays/s trace %7.1f shadow %7.1f nodes/q %.2f | flattened %7.1f" % (sys.argv[1], d["value"], s["value"], r.get("mrays_s_trace_kernel") or 0, r.get("mrays_s_shadow_kernel") or 0, r.get("nodes_per_query") or 0, f.get("value") or 0))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
