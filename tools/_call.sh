bash tools/gpu_call.sh r2ao tests bench
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ao_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'build',d['bvh_build_ms'], d['e2e'].get('bvh_build_ms'), d['e2e'].get('setup_ms'))
print('c4',d['strong_c4']['value'],d['strong_c4'].get('bvh_build_ms'),'c5',d['tile_c5']['value'],d['tile_c5'].get('bvh_build_ms'))
print('hp',d['c3_batch']['host_pointers']['mrays_s'], d['c3_batch']['host_pointers']['seconds'])
print(d['roofline'])
PY
