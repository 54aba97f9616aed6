tools/ab_variants.sh r2aq "noroot|noroot|" "root|-|" "noroot2|noroot|" "root2|-|"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2aq_pytest.log 2>&1; tail -3 gpurun_out/r2aq_pytest.log
python tools/bench_configs.py c5 --spp 32 2>&1 | tail -3
