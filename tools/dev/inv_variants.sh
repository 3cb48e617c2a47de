timeout 40 python tools/inv_variants.py 600 8 2>&1 | tail -9
MTSCOMP_B200_PARAMS=inv_persistent=1 timeout 60 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
