timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2d_qb_default.log 2>&1
MTSCOMP_B200_LIB=mtscomp_b200/_build/lib_prof.so timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2d_qb_prof.log 2>&1
grep -h "compress \|lz profile\|ratio" gpurun_out/r2d_qb_*.log
ncu --set full --clock-control none --import-source on -k regex:lz77 -s 1 -c 1 -o gpurun_out/r2_lz_d python tools/quick_bench.py 16 8 > gpurun_out/r2_ncu_lz_d.log 2>&1
