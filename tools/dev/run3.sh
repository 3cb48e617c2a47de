timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2e_qb_default.log 2>&1
grep -h "compress \|lz profile\|ratio\|decompress" gpurun_out/r2e_qb_default.log
timeout 900 python bench.py > gpurun_out/r2e_bench.log 2>&1
tail -1 gpurun_out/r2e_bench.log
