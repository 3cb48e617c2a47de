timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench120.csv python bench.py --chunks 120 --steps 1 --warmup 1 --no-legs > gpurun_out/r02_bench120_under_ncu.log 2>&1
timeout 1500 python bench.py > gpurun_out/r02_bench.log 2>&1
tail -c 1200 gpurun_out/r02_bench.log
