timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2p_qb600.log 2>&1
grep -h "decompress(own)\|exact" gpurun_out/r2p_qb600.log
