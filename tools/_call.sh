n=$(nvidia-smi -L | wc -l); echo "gpus: $n"
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi or rank_mode or cpp_host_drives" ) > gpurun_out/r2ar_pytest_multi.log 2>&1; tail -4 gpurun_out/r2ar_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 > gpurun_out/r2ar_bench_n2.json 2> gpurun_out/r2ar_bench_n2.err
tail -c 1200 gpurun_out/r2ar_bench_n2.json; tail -3 gpurun_out/r2ar_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2ar_bench_ref_n2.json 2> gpurun_out/r2ar_bench_ref_n2.err
tail -c 400 gpurun_out/r2ar_bench_ref_n2.json
