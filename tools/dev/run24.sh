timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2ac_qb600.log 2>&1
grep -h "decompress(own)\|exact" gpurun_out/r2ac_qb600.log | cut -c1-220
timeout 300 python tools/quick_bench.py 1 1 2>&1 | grep -h "decompress(own)" | tail -2
timeout 300 python tools/quick_bench.py 8 8 2>&1 | grep -h "decompress(own)" | tail -2
timeout 300 python tools/quick_bench.py 64 8 2>&1 | grep -h "decompress(own)" | tail -2
timeout 300 python tools/lfp_probe.py 2>&1 | tail -3
