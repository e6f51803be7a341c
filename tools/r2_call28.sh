timeout 900 python tools/fused_sort_timing.py 28 > gpurun_out/r2_28_fused_sort_timing.txt 2>&1; cat gpurun_out/r2_28_fused_sort_timing.txt
