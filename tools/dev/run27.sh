timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2af_qb600.log 2>&1
grep -h "decompress(ref)\|reference-written" gpurun_out/r2af_qb600.log | tail -2 | cut -c1-220
timeout 300 python tools/quick_bench.py 64 8 2>&1 | grep -h "decompress(ref)" | tail -1
timeout 300 python tools/quick_bench.py 64 8 par_lz_wide=0 2>&1 | grep -h "decompress(ref)" | tail -1
timeout 300 python tools/quick_bench.py 200 8 2>&1 | grep -h "decompress(ref)" | tail -1
timeout 300 python tools/quick_bench.py 200 8 par_lz_wide=0 2>&1 | grep -h "decompress(ref)" | tail -1
timeout 300 python tools/lfp_probe.py 2>&1 | tail -1
