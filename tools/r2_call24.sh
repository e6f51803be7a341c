set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_24_gpu_tests.txt 2>&1; tail -6 gpurun_out/r2_24_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_24_smoke.txt 2>&1; echo smoke rc=$?; grep -c "ok" gpurun_out/r2_24_smoke.txt; grep -i "error\|Traceback" gpurun_out/r2_24_smoke.txt | head -3
