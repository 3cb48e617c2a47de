set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2_qb_default.log 2>&1
for v in prof 13_13 14_13 12_12; do
  MTSCOMP_B200_LIB=mtscomp_b200/_build/lib_$v.so timeout 300 python tools/quick_bench.py 64 8 > gpurun_out/r2_qb_$v.log 2>&1
done
MTSCOMP_B200_LIB=mtscomp_b200/_build/lib_14_13.so timeout 300 python tools/quick_bench.py 64 8 lz_ctas_per_sm=1 > gpurun_out/r2_qb_14_13_1cta.log 2>&1
timeout 300 python tools/quick_bench.py 64 8 max_chain=1 > gpurun_out/r2_qb_default_mc1.log 2>&1
timeout 300 python tools/quick_bench.py 64 8 lz_ctas_per_sm=1 > gpurun_out/r2_qb_default_1cta.log 2>&1
grep -h "compress \|lz profile\|ratio" gpurun_out/r2_qb_*.log
