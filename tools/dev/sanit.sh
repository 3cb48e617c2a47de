for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_workload.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok |done" gpurun_out/r2_sanitizer_$tool.log | tail -12
done
