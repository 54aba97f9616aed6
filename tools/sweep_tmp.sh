run() { env "$@" python bench.py --steps 6 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$*', 'value %.0f'%d['value'], 'trace %.0f shadow %.0f'%(r['mrays_s_trace_kernel'], r['mrays_s_shadow_kernel']), 'nodes/q %.2f tris/q %.2f'%(r['nodes_per_query'], r['tris_per_query']), 'nodes', d['config']['bvh_nodes'], 'build %.2f ms'%d['bvh_build_ms'], {k:round(v,3) for k,v in r['kernel_ms_share'].items()})"; }
run A=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
