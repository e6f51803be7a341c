for v in "" alt168; do
  if [ -n "$v" ]; then export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_$v.so; fi
  echo "== variant '$v'"
  timeout 600 python tools/fused_sort_timing.py 28 2>&1 | grep -E "vth=0.0 sort_every=(0|1):|vth=0.3 sort_every=1:"
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "poly_mode_matches or config4 or fused_resort_matches" 2>&1 | tail -2
done
