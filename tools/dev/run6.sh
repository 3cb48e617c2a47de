timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2h_qb.log 2>&1
grep -h "compress \|ratio\|decompress\|exact" gpurun_out/r2h_qb.log
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2h_qb600.log 2>&1
grep -h "compress \|ratio\|decompress\|exact" gpurun_out/r2h_qb600.log
