set -x
ncu --set full --clock-control none --import-source on -k regex:lz77 -s 1 -c 1 -o gpurun_out/r2_lz_a python tools/quick_bench.py 16 8 > gpurun_out/r2_ncu_lz_a.log 2>&1
tail -3 gpurun_out/r2_ncu_lz_a.log
