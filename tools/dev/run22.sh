python tools/file_write_probe.py
for r in 4 2; do
timeout 600 python tools/quick_bench.py 600 8 inv_tile_runs=$r > gpurun_out/r2aa_qb600_$r.log 2>&1
echo "== inv_tile_runs=$r"; grep -h "decompress(own)\|exact" gpurun_out/r2aa_qb600_$r.log | tail -3 | cut -c1-220
done
timeout 600 python - <<'PY'
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import bench_legs, json
print(json.dumps(bench_legs.file_leg(128)))
PY
