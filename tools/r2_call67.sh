timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_67_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2_67_gpu_tests.txt
