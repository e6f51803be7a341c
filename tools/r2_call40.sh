export PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_alt168.so
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "poly or config4 or fused or determin or sorted or c2" > gpurun_out/r2_40_alt168_tests.txt 2>&1; tail -5 gpurun_out/r2_40_alt168_tests.txt; grep -E "^E  |^tests.*Error|FAILED" gpurun_out/r2_40_alt168_tests.txt | head -30
