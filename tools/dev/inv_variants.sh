timeout 60 python tools/inv_variants.py 600 8 2>&1 | tail -9
timeout 100 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
