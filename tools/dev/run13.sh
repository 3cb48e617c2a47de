timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2r_qb600.log 2>&1
grep -h "decompress\|exact" gpurun_out/r2r_qb600.log
timeout 300 python tools/quick_bench.py 1 1 > gpurun_out/r2r_qb1.log 2>&1
grep -h "decompress" gpurun_out/r2r_qb1.log | tail -3
timeout 300 python tools/lfp_probe.py > gpurun_out/r2r_lfp.log 2>&1
tail -3 gpurun_out/r2r_lfp.log
