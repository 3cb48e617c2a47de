timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for sp in 1 0; do
timeout 600 python tools/quick_bench.py 600 8 par_single_pass=$sp > gpurun_out/r2ad_qb600_$sp.log 2>&1
echo "== par_single_pass=$sp"; grep -h "decompress(ref)\|reference-written" gpurun_out/r2ad_qb600_$sp.log | tail -2 | cut -c1-220
done
timeout 300 python tools/quick_bench.py 1 1 2>&1 | grep -h "decompress(ref)" | tail -1
timeout 300 python tools/quick_bench.py 8 8 2>&1 | grep -h "decompress(ref)" | tail -1
timeout 300 python tools/lfp_probe.py 2>&1 | tail -2
