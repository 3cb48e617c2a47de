timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/quick_bench.py 600 8 2>&1 | grep -h "decompress(ref)\|reference-written" | tail -2 | cut -c1-220
