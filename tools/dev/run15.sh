timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2t_qb600.log 2>&1
grep -h "compress \|ratio\|decompress\|exact" gpurun_out/r2t_qb600.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches8.csv python tools/quick_bench.py 8 8 > gpurun_out/r2t_qb8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches1.csv python tools/quick_bench.py 1 1 > gpurun_out/r2t_qb1.log 2>&1
timeout 300 python tools/lfp_probe.py > gpurun_out/r2t_lfp.log 2>&1
tail -4 gpurun_out/r2t_lfp.log
