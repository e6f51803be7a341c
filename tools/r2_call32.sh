timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "fused or poly or sorted or config4 or c2 or checkpoint or streamed" > gpurun_out/r2_32_tests.txt 2>&1; tail -5 gpurun_out/r2_32_tests.txt
timeout 900 python tools/fused_sort_timing.py 28 > gpurun_out/r2_32_fused_sort_timing.txt 2>&1; head -20 gpurun_out/r2_32_fused_sort_timing.txt
