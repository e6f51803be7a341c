timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "fused or poly or c2 or config4" 2>&1 | tail -3
