ncu --set full --clock-control none --import-source on -k regex:lz_resolve -c 3 -o gpurun_out/r2_res_a python tools/quick_bench.py 64 8 > gpurun_out/r2_ncu_res_a.log 2>&1
tail -2 gpurun_out/r2_ncu_res_a.log
