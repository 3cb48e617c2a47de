timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/r2g_bench.log 2>&1
tail -c 6000 gpurun_out/r2g_bench.log
