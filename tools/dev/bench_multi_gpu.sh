N=${1:-2}
SC=${2:-150}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --shard-chunks $SC > gpurun_out/r02_bench_n$N.log 2> gpurun_out/r02_bench_n$N.err
tail -c 4000 gpurun_out/r02_bench_n$N.log
tail -5 gpurun_out/r02_bench_n$N.err
timeout 600 python -m pytest tests/test_gpu_sharding.py -q 2>&1 | tail -2
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
