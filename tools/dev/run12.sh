timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2q_qb600.log 2>&1
grep -h "decompress\|exact" gpurun_out/r2q_qb600.log
