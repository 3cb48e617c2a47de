timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02_bench.log 2>&1
tail -c 600 gpurun_out/r02_bench.log
