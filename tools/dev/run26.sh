for v in pb8 pb4; do
MTSCOMP_B200_LIB=mtscomp_b200/_build/lib_$v.so timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2ae_qb600_$v.log 2>&1
echo "== $v"; grep -h "decompress(ref)\|reference-written" gpurun_out/r2ae_qb600_$v.log | tail -2 | cut -c1-220
done
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2ae_qb600.log 2>&1
echo "== default (6)"; grep -h "decompress(ref)\|reference-written" gpurun_out/r2ae_qb600.log | tail -2 | cut -c1-220
