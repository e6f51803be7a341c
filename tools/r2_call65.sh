timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "fused or poly_mode_warm or determin" 2>&1 | tail -2
for d in 1 0; do
  export PICGOLF_DEFER_FIRST=$d
  echo "== PICGOLF_DEFER_FIRST=$d"
  timeout 600 python tools/fused_sort_timing.py 28 2>&1 | grep -E "vth=(0.3|1.0) sort_every=0:"
done
