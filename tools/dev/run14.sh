timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/r02_bench_b.log 2>&1
tail -c 1500 gpurun_out/r02_bench_b.log
