for v in res4 res16; do
MTSCOMP_B200_LIB=mtscomp_b200/_build/lib_$v.so timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2z_qb600_$v.log 2>&1
echo "== $v"; grep -h "decompress(own)\|own round" gpurun_out/r2z_qb600_$v.log | tail -3 | cut -c1-220
done
timeout 600 python tools/quick_bench.py 600 8 > gpurun_out/r2z_qb600.log 2>&1
echo "== default"; grep -h "decompress(own)\|own round" gpurun_out/r2z_qb600.log | tail -3 | cut -c1-220
timeout 600 python tools/latency_probe.py > gpurun_out/r2z_latency.log 2>&1
tail -12 gpurun_out/r2z_latency.log
