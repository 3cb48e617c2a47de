ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches.csv python tools/quick_bench.py 64 8 > gpurun_out/r2s_qb.log 2>&1
grep -h "compress \|decompress" gpurun_out/r2s_qb.log | head -8
