ncu --set full --clock-control none --import-source on -k regex:seg_ -c 2 -o gpurun_out/r2_seg_b python tools/quick_bench.py 16 8 > gpurun_out/r2_ncu_seg_b.log 2>&1
tail -2 gpurun_out/r2_ncu_seg_b.log
