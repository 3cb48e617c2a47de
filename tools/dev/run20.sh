timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2y_qb600.log 2>&1
grep -h "decompress\|exact" gpurun_out/r2y_qb600.log | cut -c1-220
timeout 300 python tools/quick_bench.py 1 1 2>&1 | grep -h "decompress" | tail -3
timeout 300 python tools/quick_bench.py 8 8 2>&1 | grep -h "decompress" | tail -3
timeout 300 python tools/lfp_probe.py > gpurun_out/r2y_lfp.log 2>&1
tail -3 gpurun_out/r2y_lfp.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_launches64.csv python tools/quick_bench.py 64 8 > gpurun_out/r2y_qb64.log 2>&1
