n=$(nvidia-smi -L | wc -l); echo "gpus: $n"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 > gpurun_out/r2as_bench_n8.json 2> gpurun_out/r2as_bench_n8.err
tail -c 600 gpurun_out/r2as_bench_n8.json; tail -3 gpurun_out/r2as_bench_n8.err
