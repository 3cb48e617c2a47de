timeout 600 python tools/ref_e2e_probe.py 600 1792 2048 2560 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity_holes.py tests/test_gpu_parity.py -x -q 2>&1 | tail -1
