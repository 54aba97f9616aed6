#!/bin/bash
# One GPU-box call (run through gpurun): tools/gpu_call.sh <tag> <what>...   what = smoke | tests | sanitize | anchor | bench | benchq | ref | launches | ncu_trace | ncu_shade | ncu_c4 | ncu_c4flat | multi | multi48
# Everything lands under gpurun_out/<tag>_* (the box returns at most 64 MiB per call: at most two ncu_* steps per call); summaries for profiles/ are made afterwards in the development container
# (tools/ncu_summary.py). Each step has its own timeout so that a hang cannot take the box.
tag=$1; shift
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-roofline --no-configs"
for what in "$@"; do
  case $what in
    smoke)   ( time timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $out/${tag}_smoke.log 2>&1; tail -3 $out/${tag}_smoke.log ;;
    tests)   ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest.log 2>&1; tail -8 $out/${tag}_pytest.log ;;
    sanitize) for tool in memcheck racecheck; do ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py ) > $out/${tag}_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 $out/${tag}_sanitize_$tool.log; done ;;
    anchor)  ( time timeout 600 python -m pytest tests/test_reference_anchor.py -m gpu -x -q ) > $out/${tag}_pytest_anchor.log 2>&1; tail -4 $out/${tag}_pytest_anchor.log ;;
    bench)   timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 1500 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err ;;
    benchq)  timeout 300 python bench.py --no-cpu-baseline --no-configs > $out/${tag}_benchq.json 2> $out/${tag}_benchq.err; tail -c 1200 $out/${tag}_benchq.json; tail -3 $out/${tag}_benchq.err ;;
    ref)     timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; tail -c 600 $out/${tag}_bench_ref.json ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_l.log 2>&1; tail -2 $out/${tag}_ncu_l.log ;;
    ncu_trace) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 8 --launch-count 8 -f -o $out/${tag}_prof_trace $B > $out/${tag}_ncu_t.log 2>&1; tail -2 $out/${tag}_ncu_t.log ;;
    ncu_shade) timeout 900 ncu --set full --clock-control none -k regex:'k_shadow|k_shade' --launch-skip 16 --launch-count 6 -f -o $out/${tag}_prof_shade $B > $out/${tag}_ncu_s.log 2>&1; tail -2 $out/${tag}_ncu_s.log ;;
    ncu_c4flat) CRB_FLATTEN=1 timeout 900 ncu --set full --clock-control none -k regex:k_trace --launch-skip 8 --launch-count 8 -f -o $out/${tag}_prof_c4flat python tools/bench_configs.py c4 --spp 16 > $out/${tag}_ncu_c4.log 2>&1; tail -2 $out/${tag}_ncu_c4.log ;;
    ncu_c4)  timeout 900 ncu --set full --clock-control none -k regex:k_trace2 --launch-skip 8 --launch-count 8 -f -o $out/${tag}_prof_c4 python tools/bench_configs.py c4 --spp 16 > $out/${tag}_ncu_c4tl.log 2>&1; tail -2 $out/${tag}_ncu_c4tl.log ;;
    multi)   n=$(nvidia-smi -L | wc -l)
             ( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi or rank_mode or cpp_host_drives" ) > $out/${tag}_pytest_multi.log 2>&1; tail -8 $out/${tag}_pytest_multi.log
             for g in 1 2 4 8; do [ $g -le $n ] || continue
               if [ $g -eq 1 ]; then timeout 600 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
               else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2960$g bench.py --gpus $g > $out/${tag}_bench_n$g.json 2> $out/${tag}_bench_n$g.err; fi
               tail -c 900 $out/${tag}_bench_n$g.json; tail -2 $out/${tag}_bench_n$g.err
             done ;;
    multi48) for g in 4 8; do
               timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g bench.py --gpus $g --only c4,c5 --no-roofline > $out/${tag}_bench_n$g.json 2> $out/${tag}_bench_n$g.err
               tail -c 600 $out/${tag}_bench_n$g.json; tail -2 $out/${tag}_bench_n$g.err
             done ;;
    *) echo "unknown step $what" ;;
  esac
done
