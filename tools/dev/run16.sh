timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for ob in 1 2 4 8 16; do
timeout 600 python tools/quick_bench.py 600 8 inv_order_block=$ob > gpurun_out/r2u_qb600_$ob.log 2>&1
echo "== inv_order_block=$ob"; grep -h "decompress" gpurun_out/r2u_qb600_$ob.log | tail -3 | cut -c1-220
done
grep -h "exact" gpurun_out/r2u_qb600_1.log
timeout 300 python tools/quick_bench.py 1 1 2>&1 | grep -h "decompress" | tail -3
timeout 300 python tools/quick_bench.py 8 8 2>&1 | grep -h "decompress" | tail -3
timeout 300 python tools/lfp_probe.py > gpurun_out/r2u_lfp.log 2>&1
tail -3 gpurun_out/r2u_lfp.log
ncu --set full --clock-control none --import-source on -k regex:"seg_tokens|inv_tile|par_block" -c 6 -o gpurun_out/r2u_dec python tools/quick_bench.py 16 8 > gpurun_out/r2u_ncu.log 2>&1
