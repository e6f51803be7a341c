timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_41_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2_41_gpu_tests.txt; grep -E "^E  " gpurun_out/r2_41_gpu_tests.txt | head
for v in "" old810; do
  if [ -n "$v" ]; then export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_$v.so; fi
  echo "== variant '$v'"
  timeout 600 python tools/fused_sort_timing.py 28 2>&1 | grep -E "sort_every=(0|1):"
done
