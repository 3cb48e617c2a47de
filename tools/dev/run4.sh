timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2f_qb_default.log 2>&1
grep -h "compress \|lz profile\|ratio" gpurun_out/r2f_qb_default.log
timeout 300 python tools/lfp_probe.py > gpurun_out/r2f_lfp.log 2>&1
tail -8 gpurun_out/r2f_lfp.log
